/*
 * brute.c -- TEST INFRASTRUCTURE.  An independent legal-move generator in the role shakmaty plays for the
 * reference (tests/chess.rs:555-578, fuzz/fuzz_targets/generate_moves.rs:25-40): a second opinion on the
 * oracle that shares nothing with it.  No bitboards, no tables, no attack maps: an 8x8 mailbox, pseudo-legal
 * moves found by stepping square by square, and legality decided by playing the move on a copy of the board
 * and walking outwards from the king to see whether anything attacks it (textbook chess rules).
 *
 * Input: the 112-byte position image (same layout as pabi_position_t).  Output: the legal moves in pabi's u16
 * packing (core.rs:24-66), sorted ascending, so that the caller compares SETS as the reference's differential
 * test does (it sorts both sides, tests/chess.rs:562-575).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t white[6], black[6]; /* king, queens, rooks, bishops, knights, pawns */
  uint64_t hash;
  uint16_t fullmove_counter;
  uint8_t castling, side_to_move, halfmove_clock, en_passant_square, pad[2];
} image_t;

enum { EMPTY = 0, KING = 1, QUEEN, ROOK, BISHOP, KNIGHT, PAWN }; /* + 8 for black */

typedef struct {
  int8_t sq[64];
  int stm, castling, ep;
} board_t;

static int colour_of(int piece) { return piece >> 3; }
static int kind_of(int piece) { return piece & 7; }

static void load(const image_t *p, board_t *b) {
  memset(b->sq, 0, sizeof b->sq);
  for (int k = 0; k < 6; k++)
    for (int s = 0; s < 64; s++) {
      if ((p->white[k] >> s) & 1) b->sq[s] = (int8_t)(k + 1);
      if ((p->black[k] >> s) & 1) b->sq[s] = (int8_t)(k + 1 + 8);
    }
  b->stm = p->side_to_move;
  b->castling = p->castling;
  b->ep = p->en_passant_square < 64 ? p->en_passant_square : -1;
}

static const int KNIGHT_D[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
static const int KING_D[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
static const int ROOK_D[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
static const int BISHOP_D[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};

static int on(int f, int r) { return f >= 0 && f < 8 && r >= 0 && r < 8; }

/* is `square` attacked by a piece of colour `by`? */
static int attacked(const board_t *b, int square, int by) {
  const int f = square & 7, r = square >> 3;
  for (int i = 0; i < 8; i++) {
    int nf = f + KNIGHT_D[i][0], nr = r + KNIGHT_D[i][1];
    if (on(nf, nr) && b->sq[nr * 8 + nf] == KNIGHT + 8 * by) return 1;
    nf = f + KING_D[i][0], nr = r + KING_D[i][1];
    if (on(nf, nr) && b->sq[nr * 8 + nf] == KING + 8 * by) return 1;
  }
  /* a pawn of colour `by` attacks diagonally forwards: white pawns from the rank below, black from the rank above */
  const int pr = by == 0 ? r - 1 : r + 1;
  for (int df = -1; df <= 1; df += 2)
    if (on(f + df, pr) && b->sq[pr * 8 + f + df] == PAWN + 8 * by) return 1;
  for (int d = 0; d < 4; d++) {
    for (int nf = f + ROOK_D[d][0], nr = r + ROOK_D[d][1]; on(nf, nr); nf += ROOK_D[d][0], nr += ROOK_D[d][1]) {
      const int piece = b->sq[nr * 8 + nf];
      if (piece == EMPTY) continue;
      if (colour_of(piece) == by && (kind_of(piece) == ROOK || kind_of(piece) == QUEEN)) return 1;
      break;
    }
    for (int nf = f + BISHOP_D[d][0], nr = r + BISHOP_D[d][1]; on(nf, nr); nf += BISHOP_D[d][0], nr += BISHOP_D[d][1]) {
      const int piece = b->sq[nr * 8 + nf];
      if (piece == EMPTY) continue;
      if (colour_of(piece) == by && (kind_of(piece) == BISHOP || kind_of(piece) == QUEEN)) return 1;
      break;
    }
  }
  return 0;
}

static int king_square(const board_t *b, int colour) {
  for (int s = 0; s < 64; s++)
    if (b->sq[s] == KING + 8 * colour) return s;
  return -1;
}

/* plays from->to for the side to move on a copy and reports whether its own king is then attacked */
static int leaves_king_safe(const board_t *b, int from, int to, int is_ep) {
  board_t c = *b;
  const int us = b->stm;
  c.sq[to] = c.sq[from];
  c.sq[from] = EMPTY;
  if (is_ep) c.sq[(to & 7) + (from & 56)] = EMPTY;
  return !attacked(&c, king_square(&c, us), us ^ 1);
}

static int cmp_u16(const void *a, const void *b) { return (int)*(const uint16_t *)a - (int)*(const uint16_t *)b; }

static void push(uint16_t *out, uint32_t *n, int from, int to, int promo) { out[(*n)++] = (uint16_t)(from | (to << 6) | (promo << 12)); }

uint32_t brute_legal_moves(const image_t *p, uint16_t *out /* [256] */) {
  board_t b;
  load(p, &b);
  const int us = b.stm, them = us ^ 1;
  uint32_t n = 0;
  for (int from = 0; from < 64; from++) {
    const int piece = b.sq[from];
    if (piece == EMPTY || colour_of(piece) != us) continue;
    const int f = from & 7, r = from >> 3, kind = kind_of(piece);
    if (kind == KNIGHT || kind == KING) {
      const int(*D)[2] = kind == KNIGHT ? KNIGHT_D : KING_D;
      for (int i = 0; i < 8; i++) {
        const int nf = f + D[i][0], nr = r + D[i][1];
        if (!on(nf, nr)) continue;
        const int to = nr * 8 + nf, target = b.sq[to];
        if (target != EMPTY && colour_of(target) == us) continue;
        if (leaves_king_safe(&b, from, to, 0)) push(out, &n, from, to, 0);
      }
    } else if (kind == PAWN) {
      const int dir = us == 0 ? 1 : -1, start = us == 0 ? 1 : 6, last = us == 0 ? 7 : 0;
      const int nr = r + dir;
      if (on(f, nr) && b.sq[nr * 8 + f] == EMPTY) {
        const int to = nr * 8 + f;
        if (leaves_king_safe(&b, from, to, 0)) {
          if (nr == last) for (int promo = 4; promo >= 1; promo--) push(out, &n, from, to, promo);
          else push(out, &n, from, to, 0);
        }
        if (r == start && b.sq[(nr + dir) * 8 + f] == EMPTY && leaves_king_safe(&b, from, (nr + dir) * 8 + f, 0))
          push(out, &n, from, (nr + dir) * 8 + f, 0);
      }
      for (int df = -1; df <= 1; df += 2) {
        if (!on(f + df, nr)) continue;
        const int to = nr * 8 + f + df, target = b.sq[to];
        if (target != EMPTY && colour_of(target) == them) {
          if (!leaves_king_safe(&b, from, to, 0)) continue;
          if (nr == last) for (int promo = 4; promo >= 1; promo--) push(out, &n, from, to, promo);
          else push(out, &n, from, to, 0);
        } else if (target == EMPTY && to == b.ep && leaves_king_safe(&b, from, to, 1)) {
          push(out, &n, from, to, 0);
        }
      }
    } else { /* queen, rook, bishop */
      for (int set = 0; set < 2; set++) {
        if ((set == 0 && kind == BISHOP) || (set == 1 && kind == ROOK)) continue;
        const int(*D)[2] = set == 0 ? ROOK_D : BISHOP_D;
        for (int d = 0; d < 4; d++)
          for (int nf = f + D[d][0], nr = r + D[d][1]; on(nf, nr); nf += D[d][0], nr += D[d][1]) {
            const int to = nr * 8 + nf, target = b.sq[to];
            if (target != EMPTY && colour_of(target) == us) break;
            if (leaves_king_safe(&b, from, to, 0)) push(out, &n, from, to, 0);
            if (target != EMPTY) break;
          }
      }
    }
  }
  /* castling: right held, king and rook at home, squares between them empty, the king neither in check nor
   * passing over or landing on an attacked square */
  const int home = us == 0 ? 4 : 60;
  if (b.sq[home] == KING + 8 * us && !attacked(&b, home, them)) {
    const int rights = (b.castling >> (2 * us)) & 3;
    if ((rights & 1) && b.sq[home + 3] == ROOK + 8 * us && b.sq[home + 1] == EMPTY && b.sq[home + 2] == EMPTY &&
        !attacked(&b, home + 1, them) && !attacked(&b, home + 2, them))
      push(out, &n, home, home + 2, 0);
    if ((rights & 2) && b.sq[home - 4] == ROOK + 8 * us && b.sq[home - 1] == EMPTY && b.sq[home - 2] == EMPTY &&
        b.sq[home - 3] == EMPTY && !attacked(&b, home - 1, them) && !attacked(&b, home - 2, them))
      push(out, &n, home, home - 2, 0);
  }
  qsort(out, n, sizeof(uint16_t), cmp_u16);
  return n;
}

int brute_in_check(const image_t *p) {
  board_t b;
  load(p, &b);
  return attacked(&b, king_square(&b, b.stm), b.stm ^ 1);
}

/* Compares, for n positions, the SORTED move list in `moves[offsets[i] .. offsets[i+1])` (the candidate, in any order)
 * with the brute-force list, and the candidate's in-check flags when given.  Returns the number of positions that
 * differ; first_bad receives the index of the first one. */
uint64_t brute_compare_batch(const image_t *p, uint64_t n, const uint32_t *offsets, const uint16_t *moves, const uint8_t *in_check,
                             uint64_t *first_bad, uint64_t *total_moves) {
  uint64_t bad = 0, total = 0;
  uint16_t want[256], got[256];
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t nw = brute_legal_moves(&p[i], want);
    const uint32_t ng = offsets[i + 1] - offsets[i];
    int differs = nw != ng || ng > 256;
    if (!differs) {
      memcpy(got, moves + offsets[i], ng * sizeof(uint16_t));
      qsort(got, ng, sizeof(uint16_t), cmp_u16);
      differs = memcmp(got, want, ng * sizeof(uint16_t)) != 0;
    }
    if (!differs && in_check) differs = (in_check[i] != 0) != (brute_in_check(&p[i]) != 0);
    if (differs && bad++ == 0 && first_bad) *first_bad = i;
    total += nw;
  }
  if (total_moves) *total_moves = total;
  return bad;
}

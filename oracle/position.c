/*
 * position.c -- FEN / EPD parsing, printing and validation of the CPU oracle
 * (TEST INFRASTRUCTURE; see pabi_oracle.h).
 *
 * Follows /root/reference/src/chess/position.rs:78-91 (starting), :157-275
 * (from_fen), :809-821 (TryFrom<&str>), :823-860 (Display), :934-1032
 * (validate), with the primitive parsers of core.rs and environment.rs:29-39.
 * Error texts are the reference's (tests/chess.rs:51-130 pins substrings).
 */
#include "oracle_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

static int fail(char *err, size_t cap, const char *fmt, ...) {
  if (err && cap) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, cap, fmt, ap);
    va_end(ap);
  }
  return -1;
}

static uint64_t side_all(const uint64_t s[6]) { return s[0] | s[1] | s[2] | s[3] | s[4] | s[5]; }

/* bitboard.rs:428-446 (kind priority) + position.rs:755-769 (white first).
 * Returns FEN symbol or 0. */
static char piece_at(const oracle_position_t *p, int sq) {
  static const char W[6] = {'K', 'Q', 'R', 'B', 'N', 'P'};
  static const int ORDER[6] = {ORC_KING, ORC_PAWNS, ORC_QUEENS, ORC_ROOKS, ORC_BISHOPS, ORC_KNIGHTS};
  uint64_t bit = 1ULL << sq;
  for (int side = 0; side < 2; side++) {
    const uint64_t *s = side == 0 ? p->white : p->black;
    if (!(side_all(s) & bit)) continue;
    for (int i = 0; i < 6; i++)
      if (s[ORDER[i]] & bit) return side == 0 ? W[ORDER[i]] : (char)(W[ORDER[i]] + 32);
  }
  return 0;
}

/* generated.rs:22-27 index; `kind` in core.rs:448-455 numbering. */
static uint64_t piece_key(int player, int kind, int sq) {
  return ORC_PIECE_KEYS[player * 384 + kind * 64 + sq];
}

/* position.rs:776-806 */
uint64_t oracle_compute_hash(const oracle_position_t *p) {
  static const int KIND_OF[6] = {KIND_KING, KIND_QUEEN, KIND_ROOK, KIND_BISHOP, KIND_KNIGHT, KIND_PAWN};
  uint64_t key = 0;
  if (p->side_to_move == 1) key ^= ORC_MISC_KEYS[0];
  for (int i = 0; i < 4; i++)
    if (p->castling & (1 << i)) key ^= ORC_MISC_KEYS[1 + i];
  if (p->en_passant_square < 64) key ^= ORC_EP_KEYS[p->en_passant_square & 7];
  for (int side = 0; side < 2; side++) {
    const uint64_t *s = side == 0 ? p->white : p->black;
    for (int k = 0; k < 6; k++)
      for (uint64_t b = s[k]; b; b &= b - 1) key ^= piece_key(side, KIND_OF[k], orc_lsb(b));
  }
  return key;
}

/* position.rs:78-91 */
void oracle_position_starting(oracle_position_t *out) {
  memset(out, 0, sizeof *out);
  out->white[ORC_KING] = 1ULL << 4;
  out->white[ORC_QUEENS] = 1ULL << 3;
  out->white[ORC_ROOKS] = 0x81ULL;
  out->white[ORC_BISHOPS] = 0x24ULL;
  out->white[ORC_KNIGHTS] = 0x42ULL;
  out->white[ORC_PAWNS] = 0xFF00ULL;
  for (int k = 0; k < 6; k++) out->black[k] = __builtin_bswap64(out->white[k]);
  out->castling = 15;
  out->side_to_move = 0;
  out->halfmove_clock = 0;
  out->fullmove_counter = 1;
  out->en_passant_square = 64;
  out->hash = oracle_compute_hash(out);
}

/* Rust `str::parse::<uN>()`: optional '+', then one or more ASCII digits, no
 * overflow past `max`. */
static int parse_unsigned(const char *s, size_t n, unsigned long max, unsigned long *out) {
  size_t i = 0;
  if (n > 0 && s[0] == '+') i = 1;
  if (i >= n) return -1;
  unsigned long v = 0;
  for (; i < n; i++) {
    if (s[i] < '0' || s[i] > '9') return -1;
    v = v * 10 + (unsigned long)(s[i] - '0');
    if (v > max) return -1;
  }
  *out = v;
  return 0;
}

/* core.rs:496-552 */
static int piece_from_symbol(unsigned char c, int *side, int *kind) {
  switch (c) {
    case 'P': *side = 0; *kind = ORC_PAWNS; return 0;
    case 'N': *side = 0; *kind = ORC_KNIGHTS; return 0;
    case 'B': *side = 0; *kind = ORC_BISHOPS; return 0;
    case 'R': *side = 0; *kind = ORC_ROOKS; return 0;
    case 'Q': *side = 0; *kind = ORC_QUEENS; return 0;
    case 'K': *side = 0; *kind = ORC_KING; return 0;
    case 'p': *side = 1; *kind = ORC_PAWNS; return 0;
    case 'n': *side = 1; *kind = ORC_KNIGHTS; return 0;
    case 'b': *side = 1; *kind = ORC_BISHOPS; return 0;
    case 'r': *side = 1; *kind = ORC_ROOKS; return 0;
    case 'q': *side = 1; *kind = ORC_QUEENS; return 0;
    case 'k': *side = 1; *kind = ORC_KING; return 0;
    default: return -1;
  }
}

/* core.rs:620-674: exactly the 16 ordered subsets of "KQkq", or "-". */
static int castling_from_str(const char *s, size_t n, uint8_t *out) {
  static const char *NAMES[16] = {"-",  "K",  "Q",  "KQ",  "k",  "Kk",  "Qk",  "KQk",
                                  "q",  "Kq", "Qq", "KQq", "kq", "Kkq", "Qkq", "KQkq"};
  for (int i = 0; i < 16; i++)
    if (strlen(NAMES[i]) == n && memcmp(NAMES[i], s, n) == 0) {
      *out = (uint8_t)i;
      return 0;
    }
  return -1;
}

/* Next ' '-separated part, Rust `split(' ')` semantics (empty parts kept).
 * Returns 0 when the iterator is exhausted. */
typedef struct {
  const char *cur;
  int done;
} splitter_t;

static int next_part(splitter_t *it, const char **s, size_t *n) {
  if (it->done) return 0;
  const char *sp = strchr(it->cur, ' ');
  *s = it->cur;
  if (sp) {
    *n = (size_t)(sp - it->cur);
    it->cur = sp + 1;
  } else {
    *n = strlen(it->cur);
    it->done = 1;
  }
  return 1;
}

/* position.rs:157-275 */
int oracle_position_from_fen(const char *input, oracle_position_t *out, char *err, size_t cap) {
  oracle_position_t p;
  memset(&p, 0, sizeof p);
  splitter_t parts = {input, 0};
  const char *s;
  size_t n;

  if (!next_part(&parts, &s, &n)) return fail(err, cap, "missing pieces placement");
  const char *placement = s;
  size_t placement_len = n;
  int rank_id = 8;
  size_t pos = 0;
  for (;;) { /* ranks = placement.split('/') */
    size_t end = pos;
    while (end < placement_len && placement[end] != '/') end++;
    if (rank_id == 0)
      return fail(err, cap, "expected 8 ranks, got %.*s", (int)placement_len, placement);
    rank_id -= 1;
    unsigned file = 0;
    for (size_t i = pos; i < end; i++) {
      unsigned char symbol = (unsigned char)placement[i];
      if (file > 8) return fail(err, cap, "file exceeded 8");
      if (symbol == '0') return fail(err, cap, "increment can not be 0");
      if (symbol >= '1' && symbol <= '9') {
        file += symbol - '0';
        continue;
      }
      int side, kind;
      if (piece_from_symbol(symbol, &side, &kind) != 0)
        return fail(err, cap, "piece symbol should be in \"PNBRQKpnbrqk\", got '%c'", symbol);
      if (file > 7) return fail(err, cap, "file should be within 0..BOARD_WIDTH, got %u", file);
      uint64_t bit = 1ULL << (rank_id * 8 + (int)file);
      if (side == 0) p.white[kind] |= bit; else p.black[kind] |= bit;
      file += 1;
    }
    if (file != 8)
      return fail(err, cap, "rank size should be exactly 8, got %.*s of length %u",
                  (int)(end - pos), placement + pos, file);
    if (end >= placement_len) break;
    pos = end + 1;
  }
  if (rank_id != 0)
    return fail(err, cap, "there should be 8 ranks, got %.*s", (int)placement_len, placement);

  /* environment.rs:29-39 */
  if (!next_part(&parts, &s, &n)) return fail(err, cap, "missing side to move");
  if (n == 1 && s[0] == 'w') p.side_to_move = 0;
  else if (n == 1 && s[0] == 'b') p.side_to_move = 1;
  else return fail(err, cap, "color should be 'w' or 'b', got '%.*s'", (int)n, s);

  if (!next_part(&parts, &s, &n)) return fail(err, cap, "missing castling rights");
  if (castling_from_str(s, n, &p.castling) != 0)
    return fail(err, cap, "unknown castle rights: %.*s", (int)n, s);

  if (!next_part(&parts, &s, &n)) return fail(err, cap, "missing en passant square");
  if (n == 1 && s[0] == '-') {
    p.en_passant_square = 64;
  } else { /* core.rs:286-299 */
    if (n != 2) return fail(err, cap, "square should be two-char, got %.*s with %zu chars", (int)n, s, n);
    if (s[0] < 'a' || s[0] > 'h') return fail(err, cap, "file should be within 'a'..='h', got '%c'", s[0]);
    if (s[1] < '1' || s[1] > '8') return fail(err, cap, "rank should be within '1'..='8', got '%c'", s[1]);
    p.en_passant_square = (uint8_t)((s[0] - 'a') + 8 * (s[1] - '1'));
  }

  int have_halfmove = 0, have_fullmove = 0;
  unsigned long halfmove = 0, fullmove = 1;
  if (next_part(&parts, &s, &n)) {
    if (parse_unsigned(s, n, 255, &halfmove) != 0)
      return fail(err, cap, "halfmove clock can not be parsed %.*s", (int)n, s);
    have_halfmove = 1;
  }
  if (next_part(&parts, &s, &n)) {
    if (parse_unsigned(s, n, 65535, &fullmove) != 0)
      return fail(err, cap, "fullmove counter can not be parsed %.*s", (int)n, s);
    if (fullmove == 0) return fail(err, cap, "fullmove counter can not be 0");
    have_fullmove = 1;
  } else if (have_halfmove) {
    return fail(err, cap, "if halfmove clock is present, fullmove counter must be present");
  }
  if (next_part(&parts, &s, &n)) return fail(err, cap, "trailing symbols");

  p.halfmove_clock = have_halfmove ? (uint8_t)halfmove : 0;
  p.fullmove_counter = have_fullmove ? (uint16_t)fullmove : 1;
  p.hash = oracle_compute_hash(&p);

  char why[256];
  if (oracle_validate(&p, why, sizeof why) != 0) return fail(err, cap, "illegal position: %s", why);
  *out = p;
  return 0;
}

/* position.rs:809-821: trim, strip one "fen " / "epd " prefix, from_fen. */
int oracle_position_parse(const char *text, oracle_position_t *out, char *err, size_t cap) {
  size_t len = strlen(text), a = 0, b = len;
  /* Rust str::trim() strips Unicode whitespace; the ASCII subset is what
   * the reference's tests exercise (tests/chess.rs:153-180). */
  while (a < b && (text[a] == ' ' || (text[a] >= '\t' && text[a] <= '\r'))) a++;
  while (b > a && (text[b - 1] == ' ' || (text[b - 1] >= '\t' && text[b - 1] <= '\r'))) b--;
  char buf[512];
  if (b - a >= sizeof buf) return fail(err, cap, "input too long");
  memcpy(buf, text + a, b - a);
  buf[b - a] = 0;
  const char *body = buf;
  if (strncmp(body, "fen ", 4) == 0 || strncmp(body, "epd ", 4) == 0) body += 4;
  return oracle_position_from_fen(body, out, err, cap);
}

/* position.rs:823-860 */
int oracle_position_to_fen(const oracle_position_t *p, char *out, size_t cap) {
  char buf[128];
  size_t k = 0;
  for (int rank = 7; rank >= 0; rank--) {
    int empty = 0;
    for (int file = 0; file < 8; file++) {
      char c = piece_at(p, rank * 8 + file);
      if (c) {
        if (empty) buf[k++] = (char)('0' + empty), empty = 0;
        buf[k++] = c;
      } else {
        empty++;
      }
    }
    if (empty) buf[k++] = (char)('0' + empty);
    if (rank != 0) buf[k++] = '/';
  }
  buf[k++] = ' ';
  buf[k++] = p->side_to_move == 0 ? 'w' : 'b';
  buf[k++] = ' ';
  if (p->castling == 0) buf[k++] = '-';
  if (p->castling & 1) buf[k++] = 'K';
  if (p->castling & 2) buf[k++] = 'Q';
  if (p->castling & 4) buf[k++] = 'k';
  if (p->castling & 8) buf[k++] = 'q';
  buf[k++] = ' ';
  if (p->en_passant_square < 64) {
    buf[k++] = (char)('a' + (p->en_passant_square & 7));
    buf[k++] = (char)('1' + (p->en_passant_square >> 3));
  } else {
    buf[k++] = '-';
  }
  k += (size_t)snprintf(buf + k, sizeof buf - k, " %u %u", p->halfmove_clock, p->fullmove_counter);
  if (k + 1 > cap) return -1;
  memcpy(out, buf, k + 1);
  return (int)k;
}

/* position.rs:934-1032 */
int oracle_validate(const oracle_position_t *p, char *err, size_t cap) {
  const uint64_t RANK1 = 0xFFULL, RANK8 = 0xFFULL << 56;
  if (p->fullmove_counter == 0) return fail(err, cap, "fullmove counter cannot be zero");
  if (orc_popcount(p->white[ORC_KING]) != 1)
    return fail(err, cap, "expected 1 white king, got %d", orc_popcount(p->white[ORC_KING]));
  if (orc_popcount(p->black[ORC_KING]) != 1)
    return fail(err, cap, "expected 1 black king, got %d", orc_popcount(p->black[ORC_KING]));
  if (orc_popcount(p->white[ORC_PAWNS]) > 8)
    return fail(err, cap, "expected <= 8 white pawns, got %d", orc_popcount(p->white[ORC_PAWNS]));
  if (orc_popcount(p->black[ORC_PAWNS]) > 8)
    return fail(err, cap, "expected <= 8 black pawns, got %d", orc_popcount(p->black[ORC_PAWNS]));
  if ((p->white[ORC_PAWNS] | p->black[ORC_PAWNS]) & (RANK1 | RANK8))
    return fail(err, cap, "pawns can not be placed on backranks");
  oracle_attack_info_t ai;
  oracle_attack_info(p, &ai);
  if (orc_popcount(ai.checkers) > 2)
    return fail(err, cap, "expected <= 2 checks, got %d", orc_popcount(ai.checkers));
  if (p->en_passant_square < 64) {
    int ep = p->en_passant_square;
    int us = p->side_to_move;
    const uint64_t *our = us == 0 ? p->white : p->black;
    const uint64_t *their = us == 0 ? p->black : p->white;
    int expected_rank = us == 0 ? 5 : 2; /* Rank6 / Rank3, zero based */
    if ((ep >> 3) != expected_rank)
      return fail(err, cap, "expected en passant square to be on rank %d, got %d", expected_rank + 1,
                  (ep >> 3) + 1);
    /* them pushes towards us: down if they are black. */
    int pushed_pawn = us == 0 ? ep - 8 : ep + 8;
    if (!(their[ORC_PAWNS] & (1ULL << pushed_pawn)))
      return fail(err, cap, "en passant square is not beyond pushed pawn");
    int king = orc_lsb(our[ORC_KING]);
    uint64_t occ = side_all(p->white) | side_all(p->black);
    if (ai.checkers) {
      if (orc_popcount(ai.checkers) > 1)
        return fail(err, cap, "more than 1 check after double pawn push is impossible");
      if (ai.checkers != (1ULL << pushed_pawn)) {
        int checker = orc_lsb(ai.checkers);
        int original_square = us == 0 ? ep + 8 : ep - 8;
        if (!(orc_ray(checker, king) & (1ULL << original_square)))
          return fail(err, cap,
                      "the only possible checks after double pawn push are either discovery "
                      "targeting the original pawn square or the pushed pawn itself");
      }
    }
    for (uint64_t b = their[ORC_QUEENS] | their[ORC_BISHOPS]; b; b &= b - 1) {
      int attacker = orc_lsb(b);
      uint64_t xray = orc_bishop_ray(attacker, king);
      if (orc_popcount(xray & occ) == 2 && (xray & (1ULL << attacker)) &&
          (xray & (1ULL << pushed_pawn)))
        return fail(err, cap, "doubly pushed pawn can not be the only blocker on a diagonal");
    }
  }
  return 0;
}

/* core.rs:105-123 */
int oracle_move_from_uci(const char *uci, uint16_t *out) {
  size_t n = strlen(uci);
  if (n != 4 && n != 5) return -1;
  for (int i = 0; i < 4; i += 2)
    if (uci[i] < 'a' || uci[i] > 'h' || uci[i + 1] < '1' || uci[i + 1] > '8') return -1;
  int from = (uci[0] - 'a') + 8 * (uci[1] - '1');
  int to = (uci[2] - 'a') + 8 * (uci[3] - '1');
  int promo = 0;
  if (n == 5) {
    switch (uci[4]) { /* core.rs:707-717 */
      case 'n': promo = PROMO_KNIGHT; break;
      case 'b': promo = PROMO_BISHOP; break;
      case 'r': promo = PROMO_ROOK; break;
      case 'q': promo = PROMO_QUEEN; break;
      default: return -1;
    }
  }
  *out = orc_move(from, to, promo);
  return 0;
}

/* core.rs:125-134 */
int oracle_move_to_uci(uint16_t mv, char out[6]) {
  static const char PROMO[5] = {0, 'n', 'b', 'r', 'q'};
  int from = orc_move_from(mv), to = orc_move_to(mv), promo = orc_move_promo(mv);
  out[0] = (char)('a' + (from & 7));
  out[1] = (char)('1' + (from >> 3));
  out[2] = (char)('a' + (to & 7));
  out[3] = (char)('1' + (to >> 3));
  int k = 4;
  if (promo >= 1 && promo <= 4) out[k++] = PROMO[promo];
  out[k] = 0;
  return k;
}

/*
 * movegen.c -- AttackInfo, legal move generation, make_move, in_check and perft
 * of the CPU oracle (TEST INFRASTRUCTURE; see pabi_oracle.h).
 *
 * Follows /root/reference/src/chess/attacks.rs:89-214 (AttackInfo, castle walk
 * masks) and /root/reference/src/chess/position.rs:317-408 (generate_moves),
 * :414-732 (make_move and helpers), :735-746 (in_check), :909-924 (perft),
 * :1034-1288 (per-piece generators).  Moves come out in the order the
 * reference's code emits them.
 */
#include "oracle_internal.h"

#include <string.h>

static uint64_t side_all(const uint64_t s[6]) { return s[0] | s[1] | s[2] | s[3] | s[4] | s[5]; }

/* attacks.rs:101-200 */
static void attack_info_new(int they, const uint64_t their[6], int king, uint64_t our_occupancy,
                            uint64_t occupancy, oracle_attack_info_t *r) {
  memset(r, 0, sizeof *r);
  uint64_t king_bit = 1ULL << king;
  r->safe_king_squares = ~our_occupancy & ORC_KING_ATTACKS[king];
  uint64_t occupancy_without_king = occupancy & ~king_bit;
  r->attacks |= ORC_KING_ATTACKS[orc_lsb(their[ORC_KING])];
  for (uint64_t b = their[ORC_KNIGHTS]; b; b &= b - 1) {
    int knight = orc_lsb(b);
    uint64_t targets = ORC_KNIGHT_ATTACKS[knight];
    r->attacks |= targets;
    if (targets & king_bit) r->checkers |= 1ULL << knight;
  }
  for (uint64_t b = their[ORC_PAWNS]; b; b &= b - 1) {
    int pawn = orc_lsb(b);
    uint64_t targets = orc_pawn_attacks(pawn, they);
    r->attacks |= targets;
    if (targets & king_bit) r->checkers |= 1ULL << pawn;
  }
  /* Queens, then bishops, then rooks (attacks.rs:138-197).  kind: 0 queen
   * (ray), 1 bishop (bishop_ray), 2 rook (rook_ray). */
  static const int KINDS[3] = {ORC_QUEENS, ORC_BISHOPS, ORC_ROOKS};
  for (int k = 0; k < 3; k++) {
    for (uint64_t b = their[KINDS[k]]; b; b &= b - 1) {
      int sq = orc_lsb(b);
      uint64_t targets = k == 0   ? orc_queen_attacks(sq, occupancy)
                         : k == 1 ? orc_bishop_attacks(sq, occupancy)
                                  : orc_rook_attacks(sq, occupancy);
      r->attacks |= targets;
      if (targets & king_bit) {
        r->checkers |= 1ULL << sq;
        r->safe_king_squares &= ~(k == 0   ? orc_queen_attacks(sq, occupancy_without_king)
                                  : k == 1 ? orc_bishop_attacks(sq, occupancy_without_king)
                                           : orc_rook_attacks(sq, occupancy_without_king));
        continue;
      }
      uint64_t attack_ray = k == 0   ? orc_ray(sq, king)
                            : k == 1 ? orc_bishop_ray(sq, king)
                                     : orc_rook_ray(sq, king);
      uint64_t blocker = (attack_ray & occupancy) & ~(1ULL << sq);
      if (orc_popcount(blocker) == 1) {
        if (blocker & our_occupancy) r->pins |= blocker; else r->xrays |= blocker;
      }
    }
  }
  r->safe_king_squares &= ~r->attacks;
}

/* position.rs:285-292 */
void oracle_attack_info(const oracle_position_t *p, oracle_attack_info_t *out) {
  int us = p->side_to_move;
  const uint64_t *our = us == 0 ? p->white : p->black;
  const uint64_t *their = us == 0 ? p->black : p->white;
  uint64_t our_occ = side_all(our), their_occ = side_all(their);
  attack_info_new(!us, their, orc_lsb(our[ORC_KING]), our_occ, our_occ | their_occ, out);
}

/* position.rs:735-746 */
int oracle_in_check(const oracle_position_t *p) {
  oracle_attack_info_t ai;
  oracle_attack_info(p, &ai);
  return ai.checkers != 0;
}

/* The reference repeats `pins.contains(from) && (ray(from, king) & ray(to,
 * king)).is_empty()` (position.rs:1074,1097,1128,1200,1217). */
static inline int pinned_off_line(uint64_t pins, int from, int to, int king) {
  return (pins & (1ULL << from)) && (orc_ray(from, king) & orc_ray(to, king)) == 0;
}

/* position.rs:1132-1140 / :1183-1195 */
static inline uint32_t push_pawn_move(uint16_t *moves, uint32_t n, int from, int to) {
  int rank = to >> 3;
  if (rank == 0 || rank == 7) {
    moves[n++] = orc_move(from, to, PROMO_QUEEN);
    moves[n++] = orc_move(from, to, PROMO_ROOK);
    moves[n++] = orc_move(from, to, PROMO_BISHOP);
    moves[n++] = orc_move(from, to, PROMO_KNIGHT);
  } else {
    moves[n++] = orc_move(from, to, 0);
  }
  return n;
}

/* position.rs:317-408 */
uint32_t oracle_generate_moves(const oracle_position_t *p, uint16_t moves[256]) {
  uint32_t n = 0;
  int us = p->side_to_move, them = !us;
  const uint64_t *our = us == 0 ? p->white : p->black;
  const uint64_t *their = us == 0 ? p->black : p->white;
  int king = orc_lsb(our[ORC_KING]);
  uint64_t our_occ = side_all(our), their_occ = side_all(their);
  uint64_t occ = our_occ | their_occ;
  uint64_t their_or_empty = ~our_occ;
  oracle_attack_info_t ai;
  attack_info_new(them, their, king, our_occ, occ, &ai);

  /* generate_king_moves, position.rs:1034-1040 */
  for (uint64_t b = ai.safe_king_squares; b; b &= b - 1) moves[n++] = orc_move(king, orc_lsb(b), 0);

  /* position.rs:333-358 */
  uint64_t blocking_ray;
  switch (orc_popcount(ai.checkers)) {
    case 0: blocking_ray = ~0ULL; break;
    case 1: {
      uint64_t ray = orc_ray(orc_lsb(ai.checkers), king);
      blocking_ray = ray == 0 ? ai.checkers : ray;
      break;
    }
    default: return n; /* 2; more is unreachable!() in the reference */
  }

  /* generate_knight_moves, position.rs:1042-1059 */
  for (uint64_t b = our[ORC_KNIGHTS] & ~ai.pins; b; b &= b - 1) {
    int from = orc_lsb(b);
    for (uint64_t t = ORC_KNIGHT_ATTACKS[from] & their_or_empty & blocking_ray; t; t &= t - 1)
      moves[n++] = orc_move(from, orc_lsb(t), 0);
  }
  /* generate_rook_moves over rooks|queens, position.rs:1061-1081 */
  for (uint64_t b = our[ORC_ROOKS] | our[ORC_QUEENS]; b; b &= b - 1) {
    int from = orc_lsb(b);
    for (uint64_t t = orc_rook_attacks(from, occ) & their_or_empty & blocking_ray; t; t &= t - 1) {
      int to = orc_lsb(t);
      if (pinned_off_line(ai.pins, from, to, king)) continue;
      moves[n++] = orc_move(from, to, 0);
    }
  }
  /* generate_bishop_moves over bishops|queens, position.rs:1083-1104 */
  for (uint64_t b = our[ORC_BISHOPS] | our[ORC_QUEENS]; b; b &= b - 1) {
    int from = orc_lsb(b);
    for (uint64_t t = orc_bishop_attacks(from, occ) & their_or_empty & blocking_ray; t; t &= t - 1) {
      int to = orc_lsb(t);
      if (pinned_off_line(ai.pins, from, to, king)) continue;
      moves[n++] = orc_move(from, to, 0);
    }
  }

  /* generate_pawn_moves, position.rs:1106-1224 */
  uint64_t pawns = our[ORC_PAWNS];
  for (uint64_t b = pawns; b; b &= b - 1) { /* captures, :1123-1142 */
    int from = orc_lsb(b);
    for (uint64_t t = orc_pawn_attacks(from, us) & their_occ & their_or_empty & blocking_ray; t;
         t &= t - 1) {
      int to = orc_lsb(t);
      if (pinned_off_line(ai.pins, from, to, king)) continue;
      n = push_pawn_move(moves, n, from, to);
    }
  }
  if (p->en_passant_square < 64) { /* :1144-1178 */
    int ep = p->en_passant_square;
    /* ep.shift(pawn_push_direction(them)): them = white pushes Up (+8). */
    int ep_pawn = them == 0 ? ep + 8 : ep - 8;
    uint64_t candidates = orc_pawn_attacks(ep, them) & pawns;
    if (ep_pawn >= 0 && ep_pawn < 64 && (ai.checkers & (1ULL << ep_pawn))) {
      for (uint64_t b = candidates; b; b &= b - 1) {
        int our_pawn = orc_lsb(b);
        if (ai.pins & (1ULL << our_pawn)) continue;
        moves[n++] = orc_move(our_pawn, ep, 0);
      }
    } else {
      for (uint64_t b = candidates; b; b &= b - 1) {
        int our_pawn = orc_lsb(b);
        uint64_t after = occ;
        after &= ~(1ULL << our_pawn);
        if (ep_pawn >= 0 && ep_pawn < 64) after &= ~(1ULL << ep_pawn);
        after |= 1ULL << ep;
        if ((orc_queen_attacks(king, after) & their[ORC_QUEENS]) == 0 &&
            (orc_rook_attacks(king, after) & their[ORC_ROOKS]) == 0 &&
            (orc_bishop_attacks(king, after) & their[ORC_BISHOPS]) == 0)
          moves[n++] = orc_move(our_pawn, ep, 0);
      }
    }
  }
  /* pushes, :1179-1223; shift(Up) = << 8, shift(Down) = >> 8 (bitboard.rs:124-129) */
  uint64_t pawn_pushes = (us == 0 ? pawns << 8 : pawns >> 8) & ~occ;
  uint64_t original_squares = us == 0 ? pawn_pushes >> 8 : pawn_pushes << 8;
  for (uint64_t f = original_squares, t = pawn_pushes; f && t; f &= f - 1, t &= t - 1) {
    int from = orc_lsb(f), to = orc_lsb(t);
    if (!(blocking_ray & (1ULL << to))) continue;
    if (pinned_off_line(ai.pins, from, to, king)) continue;
    n = push_pawn_move(moves, n, from, to);
  }
  uint64_t start_rank = us == 0 ? 0xFF00ULL : 0x00FF000000000000ULL; /* core.rs:407-412 */
  uint64_t third_rank = us == 0 ? start_rank << 8 : start_rank >> 8;
  uint64_t double_pushes = (us == 0 ? (pawn_pushes & third_rank) << 8 : (pawn_pushes & third_rank) >> 8) & ~occ;
  uint64_t double_origins = us == 0 ? double_pushes >> 16 : double_pushes << 16;
  for (uint64_t f = double_origins, t = double_pushes; f && t; f &= f - 1, t &= t - 1) {
    int from = orc_lsb(f), to = orc_lsb(t);
    if (!(blocking_ray & (1ULL << to))) continue;
    if (pinned_off_line(ai.pins, from, to, king)) continue;
    moves[n++] = orc_move(from, to, 0);
  }

  /* generate_castle_moves, position.rs:1226-1288; masks attacks.rs:203-214 */
  if (ai.checkers == 0) {
    if (us == 0) {
      if ((p->castling & 1) && (ai.attacks & 0x60ULL) == 0 && (occ & (0x60ULL | 0x60ULL)) == 0)
        moves[n++] = orc_move(4, 6, 0);
      if ((p->castling & 2) && (ai.attacks & 0x0CULL) == 0 && (occ & (0x0CULL | 0x0EULL)) == 0)
        moves[n++] = orc_move(4, 2, 0);
    } else {
      if ((p->castling & 4) && (ai.attacks & (0x60ULL << 56)) == 0 && (occ & (0x60ULL << 56)) == 0)
        moves[n++] = orc_move(60, 62, 0);
      if ((p->castling & 8) && (ai.attacks & (0x0CULL << 56)) == 0 && (occ & (0x0EULL << 56)) == 0)
        moves[n++] = orc_move(60, 58, 0);
    }
  }
  return n;
}

static inline uint64_t piece_key(int player, int kind, int sq) {
  return ORC_PIECE_KEYS[player * 384 + kind * 64 + sq];
}

/* position.rs:414-433 with helpers :435-732 */
void oracle_make_move(oracle_position_t *p, uint16_t mv) {
  int from = orc_move_from(mv), to = orc_move_to(mv), promo = orc_move_promo(mv);
  uint64_t from_bit = 1ULL << from, to_bit = 1ULL << to;
  int us = p->side_to_move;
  uint64_t *our = us == 0 ? p->white : p->black;
  uint64_t *their = us == 0 ? p->black : p->white;

  p->halfmove_clock = (uint8_t)(p->halfmove_clock + 1); /* :419 */

  /* update_castling_rights, :435-468 (E1 4, H1 7, A1 0, E8 60, H8 63, A8 56) */
  if ((p->castling & 1) && (from == 4 || from == 7 || to == 7)) p->castling &= ~1, p->hash ^= ORC_MISC_KEYS[1];
  if ((p->castling & 2) && (from == 4 || from == 0 || to == 0)) p->castling &= ~2, p->hash ^= ORC_MISC_KEYS[2];
  if ((p->castling & 4) && (from == 60 || from == 63 || to == 63)) p->castling &= ~4, p->hash ^= ORC_MISC_KEYS[3];
  if ((p->castling & 8) && (from == 60 || from == 56 || to == 56)) p->castling &= ~8, p->hash ^= ORC_MISC_KEYS[4];

  /* handle_capture, :470-502 (queens, rooks, bishops, knights, pawns) */
  if (side_all(their) & to_bit) {
    static const int ORDER[5] = {ORC_QUEENS, ORC_ROOKS, ORC_BISHOPS, ORC_KNIGHTS, ORC_PAWNS};
    static const int KIND[5] = {KIND_QUEEN, KIND_ROOK, KIND_BISHOP, KIND_KNIGHT, KIND_PAWN};
    p->halfmove_clock = 0;
    for (int i = 0; i < 5; i++)
      if (their[ORDER[i]] & to_bit) {
        their[ORDER[i]] &= ~to_bit;
        p->hash ^= piece_key(!us, KIND[i], to);
        break;
      }
  }

  /* make_pawn_move, :504-618 */
  int previous_ep = p->en_passant_square;
  p->en_passant_square = 64;
  if (our[ORC_PAWNS] & from_bit) {
    p->halfmove_clock = 0;
    if (previous_ep < 64 && to == previous_ep) {
      int captured = (to & 7) + 8 * (from >> 3);
      their[ORC_PAWNS] &= ~(1ULL << captured);
      p->hash ^= piece_key(!us, KIND_PAWN, captured);
    }
    our[ORC_PAWNS] &= ~from_bit;
    p->hash ^= piece_key(us, KIND_PAWN, from);
    if (promo != 0) {
      switch (promo) {
        case PROMO_QUEEN: our[ORC_QUEENS] |= to_bit; p->hash ^= piece_key(us, KIND_QUEEN, to); break;
        case PROMO_ROOK: our[ORC_ROOKS] |= to_bit; p->hash ^= piece_key(us, KIND_ROOK, to); break;
        case PROMO_BISHOP: our[ORC_BISHOPS] |= to_bit; p->hash ^= piece_key(us, KIND_BISHOP, to); break;
        default: our[ORC_KNIGHTS] |= to_bit; p->hash ^= piece_key(us, KIND_KNIGHT, to); break;
      }
    } else {
      our[ORC_PAWNS] |= to_bit;
      p->hash ^= piece_key(us, KIND_PAWN, to);
      int single_push_square = us == 0 ? from + 8 : from - 8;
      int start_rank = us == 0 ? 1 : 6;
      /* :606-615: en passant square only if an enemy pawn attacks it. */
      if ((from >> 3) == start_rank && (from & 7) == (to & 7) && single_push_square != to &&
          (their[ORC_PAWNS] & orc_pawn_attacks(single_push_square, us)) != 0) {
        p->en_passant_square = (uint8_t)single_push_square;
        p->hash ^= ORC_EP_KEYS[single_push_square & 7];
      }
    }
  }

  /* make_king_move, :622-698 */
  if (our[ORC_KING] & from_bit) {
    int backrank = us == 0 ? 0 : 7;
    if ((from >> 3) == backrank && (to >> 3) == backrank && (from & 7) == 4) {
      if ((to & 7) == 6) {
        int rf = 7 + 8 * backrank, rt = 5 + 8 * backrank;
        our[ORC_ROOKS] &= ~(1ULL << rf);
        p->hash ^= piece_key(us, KIND_ROOK, rf);
        our[ORC_ROOKS] |= 1ULL << rt;
        p->hash ^= piece_key(us, KIND_ROOK, rt);
      } else if ((to & 7) == 2) {
        int rf = 0 + 8 * backrank, rt = 3 + 8 * backrank;
        our[ORC_ROOKS] &= ~(1ULL << rf);
        p->hash ^= piece_key(us, KIND_ROOK, rf);
        our[ORC_ROOKS] |= 1ULL << rt;
        p->hash ^= piece_key(us, KIND_ROOK, rt);
      }
    }
    our[ORC_KING] &= ~from_bit;
    p->hash ^= piece_key(us, KIND_KING, from);
    our[ORC_KING] |= to_bit;
    p->hash ^= piece_key(us, KIND_KING, to);
  }

  /* make_regular_move, :700-732 (queens, rooks, bishops, knights) */
  {
    static const int ORDER[4] = {ORC_QUEENS, ORC_ROOKS, ORC_BISHOPS, ORC_KNIGHTS};
    static const int KIND[4] = {KIND_QUEEN, KIND_ROOK, KIND_BISHOP, KIND_KNIGHT};
    for (int i = 0; i < 4; i++)
      if (our[ORDER[i]] & from_bit) {
        our[ORDER[i]] &= ~from_bit;
        p->hash ^= piece_key(us, KIND[i], from);
        our[ORDER[i]] |= to_bit;
        p->hash ^= piece_key(us, KIND[i], to);
        break;
      }
  }

  if (us == 1) p->fullmove_counter = (uint16_t)(p->fullmove_counter + 1); /* :428-430 */
  p->side_to_move = (uint8_t)!us;                                          /* :432 */
}

/* position.rs:909-924 */
uint64_t oracle_perft(const oracle_position_t *p, uint8_t depth) {
  if (depth == 0) return 1;
  uint16_t moves[256];
  uint32_t n = oracle_generate_moves(p, moves);
  if (depth == 1) return n;
  uint64_t nodes = 0;
  for (uint32_t i = 0; i < n; i++) {
    oracle_position_t next = *p;
    oracle_make_move(&next, moves[i]);
    nodes += oracle_perft(&next, (uint8_t)(depth - 1));
  }
  return nodes;
}

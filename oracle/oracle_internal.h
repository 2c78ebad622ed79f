/* oracle_internal.h -- shared inline helpers of the CPU oracle (TEST
 * INFRASTRUCTURE; see pabi_oracle.h). */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H

#include "pabi_oracle.h"

#ifdef __BMI2__
#include <immintrin.h>
#endif

#define ORC_BISHOP_ATTACKS_COUNT 5248  /* generated.rs:30 */
#define ORC_ROOK_ATTACKS_COUNT 102400 /* generated.rs:44 */

extern uint64_t ORC_KNIGHT_ATTACKS[64], ORC_KING_ATTACKS[64];
extern uint64_t ORC_WHITE_PAWN_ATTACKS[64], ORC_BLACK_PAWN_ATTACKS[64];
extern uint64_t ORC_BISHOP_MASKS[64], ORC_ROOK_MASKS[64];
extern uint64_t ORC_BISHOP_OFFSETS[64], ORC_ROOK_OFFSETS[64];
extern uint64_t ORC_BISHOP_ATTACKS[ORC_BISHOP_ATTACKS_COUNT];
extern uint64_t ORC_ROOK_ATTACKS[ORC_ROOK_ATTACKS_COUNT];
extern uint64_t ORC_RAYS[4096], ORC_BISHOP_RAYS[4096], ORC_ROOK_RAYS[4096];
/* [player*384 + kind*64 + sq] with kind in core.rs:448-455 order (generated.rs:22-27). */
extern uint64_t ORC_PIECE_KEYS[768], ORC_EP_KEYS[8];
/* black to move, white short, white long, black short, black long (generated.rs:7-13). */
extern uint64_t ORC_MISC_KEYS[5];

void orc_init_tables(void);

/* PieceKind numbering used by the Zobrist key index (core.rs:448-455). */
enum { KIND_PAWN = 0, KIND_KNIGHT, KIND_BISHOP, KIND_ROOK, KIND_QUEEN, KIND_KING };

static inline int orc_popcount(uint64_t b) { return __builtin_popcountll(b); }
static inline int orc_lsb(uint64_t b) { return __builtin_ctzll(b); }

/* attacks.rs:67-87 */
static inline uint64_t orc_pext(uint64_t a, uint64_t mask) {
#ifdef __BMI2__
  return _pext_u64(a, mask);
#else
  uint64_t result = 0, scanning_bit = 1;
  while (mask != 0) {
    uint64_t ls1b = 1ULL << orc_lsb(mask);
    if (a & ls1b) result |= scanning_bit;
    mask ^= ls1b;
    scanning_bit <<= 1;
  }
  return result;
#endif
}

/* attacks.rs:27-33 */
static inline uint64_t orc_rook_attacks(int sq, uint64_t occ) {
  return ORC_ROOK_ATTACKS[ORC_ROOK_OFFSETS[sq] + orc_pext(occ, ORC_ROOK_MASKS[sq])];
}
/* attacks.rs:35-41 */
static inline uint64_t orc_bishop_attacks(int sq, uint64_t occ) {
  return ORC_BISHOP_ATTACKS[ORC_BISHOP_OFFSETS[sq] + orc_pext(occ, ORC_BISHOP_MASKS[sq])];
}
/* attacks.rs:23-25 */
static inline uint64_t orc_queen_attacks(int sq, uint64_t occ) {
  return orc_bishop_attacks(sq, occ) | orc_rook_attacks(sq, occ);
}
/* attacks.rs:47-52; player 0 = white */
static inline uint64_t orc_pawn_attacks(int sq, int player) {
  return player == 0 ? ORC_WHITE_PAWN_ATTACKS[sq] : ORC_BLACK_PAWN_ATTACKS[sq];
}
/* attacks.rs:54-64 */
static inline uint64_t orc_ray(int from, int to) { return ORC_RAYS[from * 64 + to]; }
static inline uint64_t orc_bishop_ray(int from, int to) { return ORC_BISHOP_RAYS[from * 64 + to]; }
static inline uint64_t orc_rook_ray(int from, int to) { return ORC_ROOK_RAYS[from * 64 + to]; }

/* Move packing (core.rs:24-66): bits 0-5 from, 6-11 to, 12-14 promotion
 * (1 N, 2 B, 3 R, 4 Q; core.rs:700-705). */
#define PROMO_KNIGHT 1
#define PROMO_BISHOP 2
#define PROMO_ROOK 3
#define PROMO_QUEEN 4
static inline uint16_t orc_move(int from, int to, int promo) {
  return (uint16_t)(from | (to << 6) | (promo << 12));
}
static inline int orc_move_from(uint16_t m) { return m & 63; }
static inline int orc_move_to(uint16_t m) { return (m >> 6) & 63; }
static inline int orc_move_promo(uint16_t m) { return (m >> 12) & 7; }

#endif

/*
 * chess.cuh -- the per-position arithmetic of the B200 perft / move-generation
 * path: compact position record, attack analysis, counting and emitting move
 * generators, make_move.
 *
 * Everything here is `__host__ __device__` so the same source is (a) inlined
 * into the sm_100a kernels of kernels.cu and (b) compiled for the host by
 * tests/native/hostcheck.cpp, which lets the CPU test-suite compare this exact
 * logic with the oracle before any GPU time is spent.  The shipped library only
 * ever calls it from kernels.
 *
 * Behaviour follows pabi (file:line = /root/reference/src/chess/...):
 *   attack analysis   attacks.rs:101-200   (AttackInfo::new)
 *   generate_moves    position.rs:317-408  (+ per-piece generators :1034-1288)
 *   make_move         position.rs:414-732
 *   Move packing      core.rs:24-66, promotions core.rs:700-705
 * but is organised for a GPU thread: one 64-byte record per position, set-wise
 * pawn generation, pins found by x-raying from the king, no move objects when
 * only the count is needed (pabi's perft bulk-counts at depth 1,
 * position.rs:914-916).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline __attribute__((always_inline))
#endif

namespace pabi {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

/* ---- bit helpers ------------------------------------------------------- */
PB_HD int popc(u64 x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
/* index of the least significant set bit; x != 0 (bitboard.rs:310-322 iterates LS1B first) */
PB_HD int lsb(u64 x) {
#ifdef __CUDA_ARCH__
  const u32 lo = (u32)x, hi = (u32)(x >> 32); /* 6 SASS instructions instead of __ffsll's 10 */
  return (lo ? 0 : 32) + __ffs((int)(lo ? lo : hi)) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
PB_HD u64 bit(int sq) { return 1ULL << sq; }

/* Iterates the squares of a bitboard in ascending order (BitboardIterator, bitboard.rs:303-329),
 * one 32-bit word at a time: per step a 32-bit find-first-set and clear instead of their 64-bit
 * forms, which matters on a 32-bit datapath whose ALU pipe is the bottleneck. */
struct Squares {
  u32 word, high;
  int base;
  PB_HD explicit Squares(u64 b) : word((u32)b), high((u32)(b >> 32)), base(0) {
    if (!word) word = high, high = 0, base = 32;
  }
  PB_HD bool any() const { return word != 0; }
  PB_HD int pop() {
#ifdef __CUDA_ARCH__
    const int sq = base + __ffs((int)word) - 1;
#else
    const int sq = base + __builtin_ctz(word);
#endif
    word &= word - 1;
    if (!word) word = high, high = 0, base = 32;
    return sq;
  }
};

/* ---- constants --------------------------------------------------------- */
constexpr u64 FILE_A = 0x0101010101010101ULL;
constexpr u64 FILE_H = 0x8080808080808080ULL;
constexpr u64 RANK_1 = 0x00000000000000FFULL;
constexpr u64 RANK_2 = 0x000000000000FF00ULL;
constexpr u64 RANK_3 = 0x0000000000FF0000ULL;
constexpr u64 RANK_6 = 0x0000FF0000000000ULL;
constexpr u64 RANK_8 = 0xFF00000000000000ULL;
constexpr int NO_SQUARE = 64;

/* Move packing, core.rs:24-66: bits 0-5 from, 6-11 to, 12-14 promotion
 * (0 none, 1 N, 2 B, 3 R, 4 Q; core.rs:700-705). */
constexpr int PROMO_N = 1, PROMO_B = 2, PROMO_R = 3, PROMO_Q = 4;
/* kind of the moving piece, passed along with a generated move so that make_move_kind does
 * not have to look the piece up (numbering of PieceKind, core.rs:448-455) */
constexpr int KIND_PAWN = 0, KIND_KNIGHT = 1, KIND_BISHOP = 2, KIND_ROOK = 3, KIND_QUEEN = 4, KIND_KING = 5;
PB_HD u16 pack_move(int from, int to, int promo) { return (u16)(from | (to << 6) | (promo << 12)); }

/* ---- the 64-byte device record --------------------------
 * word 0 white occupancy, words 1..6 pawns, knights, bishops, rooks, queens,
 * kings (both colours merged), word 7 meta:
 *   bits 0-3 castling (core.rs:604-612: 1 K, 2 Q, 4 k, 8 q), bits 4-10 en
 *   passant square (64 = none), bit 11 side to move (0 white), bits 16-23
 *   halfmove clock, bits 32-47 fullmove counter.
 * In HBM the records are stored as four arrays of word pairs (load_pos / store_pos below): pair k of record i lives
 * at base[(k * stride + i) * 2], so a warp reads/writes 512 contiguous bytes per 128-bit access. */
struct Pos {
  u64 white, pawns, knights, bishops, rooks, queens, kings, meta;
};
constexpr int POS_WORDS = 8;

PB_HD int meta_castling(u64 m) { return (int)(m & 15); }
PB_HD int meta_ep(u64 m) { return (int)((m >> 4) & 127); }
PB_HD int meta_stm(u64 m) { return (int)((m >> 11) & 1); }
PB_HD int meta_halfmove(u64 m) { return (int)((m >> 16) & 255); }
PB_HD int meta_fullmove(u64 m) { return (int)((m >> 32) & 0xFFFF); }
PB_HD u64 make_meta(int castling, int ep, int stm, int halfmove, int fullmove) {
  return (u64)castling | ((u64)ep << 4) | ((u64)stm << 11) | ((u64)halfmove << 16) |
         ((u64)fullmove << 32);
}

/* -DPABI_BOUNDS (a checking build, `make bounds`): every record access and every shared-memory pool access of the
 * kernels tests its index; a violation records its source line (the first one wins) and the access is skipped.
 * compute-sanitizer is not available on this GPU pool, so the GPU test-suite is run once against this build instead
 * (tools/bounds_suite.sh); pabi_cuda_bounds_violation() reads the line back. */
#if defined(PABI_BOUNDS) && defined(__CUDACC__)
__device__ unsigned int g_bounds_violation = 0;
#endif
#if defined(PABI_BOUNDS) && defined(__CUDA_ARCH__)
__device__ __forceinline__ bool in_bounds(unsigned long long i, unsigned long long n, int line) {
  if (i < n) return true;
  atomicCAS(&g_bounds_violation, 0u, (unsigned int)line);
  return false;
}
#define PB_IN_BOUNDS(i, n) pabi::in_bounds((unsigned long long)(i), (unsigned long long)(n), __LINE__)
#else
#define PB_IN_BOUNDS(i, n) true
#endif

/* Records in HBM: the eight words in PAIRS, pair k of record i at base[(k * stride + i) * 2 .. + 2): a record moves with four
 * 128-bit accesses, a warp reads or writes 512 contiguous bytes per access.  (Round 1 kept eight separate word arrays; the
 * paired arrays took the breadth-first expansion K1 from 2.04 to 1.77 ms per perft(8): half as many load/store
 * instructions for the same bytes.) */
PB_HD Pos load_pos(const u64* __restrict__ base, size_t stride, size_t i) {
  Pos p;
  if (!PB_IN_BOUNDS(i, stride)) i = 0;
#ifdef __CUDA_ARCH__
  const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(base);
  const ulonglong2 a = b2[0 * stride + i], b = b2[1 * stride + i], c = b2[2 * stride + i], d = b2[3 * stride + i];
  p.white = a.x, p.pawns = a.y, p.knights = b.x, p.bishops = b.y, p.rooks = c.x, p.queens = c.y, p.kings = d.x, p.meta = d.y;
#else
  p.white = base[(0 * stride + i) * 2], p.pawns = base[(0 * stride + i) * 2 + 1];
  p.knights = base[(1 * stride + i) * 2], p.bishops = base[(1 * stride + i) * 2 + 1];
  p.rooks = base[(2 * stride + i) * 2], p.queens = base[(2 * stride + i) * 2 + 1];
  p.kings = base[(3 * stride + i) * 2], p.meta = base[(3 * stride + i) * 2 + 1];
#endif
  return p;
}
PB_HD void store_pos(u64* __restrict__ base, size_t stride, size_t i, const Pos& p) {
  if (!PB_IN_BOUNDS(i, stride)) return;
#ifdef __CUDA_ARCH__
  ulonglong2* b2 = reinterpret_cast<ulonglong2*>(base);
  b2[0 * stride + i] = make_ulonglong2(p.white, p.pawns);
  b2[1 * stride + i] = make_ulonglong2(p.knights, p.bishops);
  b2[2 * stride + i] = make_ulonglong2(p.rooks, p.queens);
  b2[3 * stride + i] = make_ulonglong2(p.kings, p.meta);
#else
  base[(0 * stride + i) * 2] = p.white, base[(0 * stride + i) * 2 + 1] = p.pawns;
  base[(1 * stride + i) * 2] = p.knights, base[(1 * stride + i) * 2 + 1] = p.bishops;
  base[(2 * stride + i) * 2] = p.rooks, base[(2 * stride + i) * 2 + 1] = p.queens;
  base[(3 * stride + i) * 2] = p.kings, base[(3 * stride + i) * 2 + 1] = p.meta;
#endif
}

/* ---- tables ------------------------------------------------------------
 * Same contents as pabi's generated tables (generated.rs:30-82), regenerated
 * from their definition on the host (tables.cpp):
 *  - leaper tables and the bishop attack table (pabi's 5 248 entries, per-square
 *    blocks at pabi's offsets; inside a block the entries are addressed by a
 *    multiply-shift hash instead of PEXT, which sm_100a does not have) live in
 *    shared memory, staged once per block;
 *  - the rook attack table (pabi's 102 400 entries, 800 KB) and RAYS (32 KB) stay
 *    in global memory and are read through L1 (ld.global.nc).
 * `SliderEntry` is one 16-byte shared-memory load. */
struct SliderEntry {
  u64 mask;  /* relevant occupancy, generated/..._relevant_occupancies.rs */
  u64 magic; /* index = offset + (((occ & mask) * magic) >> (64 - bits)) */
};
struct SliderIndex {
  u32 offset; /* start of the square's block, generated/..._attack_offsets.rs */
  u32 shift;  /* 32 - bits: applied to the high word of the product */
};
constexpr int BISHOP_TABLE_SIZE = 5248;  /* generated.rs:30 */
constexpr int ROOK_TABLE_SIZE = 102400; /* generated.rs:44 */

struct SharedTables {
  u64 knight[64];
  u64 king[64];
  u64 diag[64]; /* full a1-h8-direction line through the square (square included) */
  u64 anti[64]; /* full a8-h1-direction line through the square */
  u64 orth[64]; /* the square's rank and file (a table load is cheaper than two variable 64-bit shifts on
                   the ALU pipe, which is the pipe this code saturates) */
  SliderEntry bishop[64];
  SliderEntry rook[64];
  SliderIndex bishop_os[64];
  SliderIndex rook_os[64];
  /* for a (white-normalised) king on the square: where an enemy knight / diagonal slider /
   * orthogonal slider must stand to reach any square next to the king or on its castling walk */
  u64 knight_zone[64];
  u64 diag_zone[64];
  u64 orth_zone[64];
  u8 keep_from[64]; /* castling rights kept when a move leaves this square (position.rs:435-468) */
  u8 keep_to[64];   /* ... when a move lands on this square */
  u64 bishop_attacks[BISHOP_TABLE_SIZE];
};

struct GlobalTables {
  const u64* rook_attacks; /* [ROOK_TABLE_SIZE] */
  const u64* rays;         /* [64*64], attacks.rs:54-56: [from, to) on a common line, else 0 */
  const u64* zobrist;      /* [ZOBRIST_KEYS], see below; read only by the paths that maintain Position::hash */
};

/* Zobrist keys (generated.rs:7-27).  Layout of the key array:
 *   [0, 768)   piece keys, index player * 384 + kind * 64 + square (get_piece_key, generated.rs:22-27;
 *              kind in PieceKind order, core.rs:448-455)
 *   [768, 776) en passant file keys (EN_PASSANT_FILES)
 *   776        BLACK_TO_MOVE
 *   777..780   WHITE_CAN_CASTLE_SHORT, WHITE_CAN_CASTLE_LONG, BLACK_CAN_CASTLE_SHORT, BLACK_CAN_CASTLE_LONG
 *              (bit order of CastleRights, core.rs:604-612)
 * The last five are the reference's fixed constants (generated.rs:7-13).  The reference draws the piece and en
 * passant keys at random on every build (build.rs:26-39), so their values cannot be matched; here they are FIXED:
 * consecutive outputs of splitmix64 seeded with 0x70616269 ("pabi"), 768 piece keys then 8 file keys
 * (tables.cpp; the oracle uses the same documented recipe). */
constexpr int ZOBRIST_EP = 768, ZOBRIST_BLACK_TO_MOVE = 776, ZOBRIST_CASTLE = 777, ZOBRIST_KEYS = 781;
constexpr u64 ZOBRIST_SEED = 0x70616269ULL;

PB_HD u64 ldg64(const u64* p) {
#ifdef __CUDA_ARCH__
  return __ldg((const unsigned long long*)p);
#else
  return *p;
#endif
}

struct Tables {
  const SharedTables* s;
  GlobalTables g;

  PB_HD u64 knight(int sq) const { return s->knight[sq]; }
  PB_HD u64 king(int sq) const { return s->king[sq]; }
  /* attacks.rs:35-41 */
  PB_HD u64 bishop(int sq, u64 occ) const {
    const SliderEntry e = s->bishop[sq];
    const SliderIndex os = s->bishop_os[sq];
    const u32 h = (u32)(((occ & e.mask) * e.magic) >> 32);
    return s->bishop_attacks[os.offset + (h >> os.shift)];
  }
  /* attacks.rs:27-33 */
  PB_HD u64 rook(int sq, u64 occ) const {
    const SliderEntry e = s->rook[sq];
    const SliderIndex os = s->rook_os[sq];
    const u32 h = (u32)(((occ & e.mask) * e.magic) >> 32);
    return ldg64(g.rook_attacks + os.offset + (h >> os.shift));
  }
  /* attacks.rs:54-56 */
  PB_HD u64 ray(int from, int to) const { return ldg64(g.rays + from * 64 + to); }
  /* The full line through `king` and `sq` (they are known to be aligned). */
  PB_HD u64 line(int king, int sq) const {
    if ((king >> 3) == (sq >> 3)) return RANK_1 << (king & 56);
    if ((king & 7) == (sq & 7)) return FILE_A << (king & 7);
    u64 d = s->diag[king];
    return ((d >> sq) & 1) ? d : s->anti[king];
  }
};

/* Pawn capture targets of a set of pawns (generated white/black_pawn_attacks.rs
 * hold the same per-square sets; pawns never stand on ranks 1/8). */
PB_HD u64 pawn_attacks_west(u64 pawns, int colour) { /* towards file a */
  return colour == 0 ? (pawns & ~FILE_A) << 7 : (pawns & ~FILE_A) >> 9;
}
PB_HD u64 pawn_attacks_east(u64 pawns, int colour) { /* towards file h */
  return colour == 0 ? (pawns & ~FILE_H) << 9 : (pawns & ~FILE_H) >> 7;
}
PB_HD u64 pawn_attacks(u64 pawns, int colour) {
  return pawn_attacks_west(pawns, colour) | pawn_attacks_east(pawns, colour);
}
PB_HD u64 push(u64 b, int colour) { return colour == 0 ? b << 8 : b >> 8; }
PB_HD u64 unpush(u64 b, int colour) { return colour == 0 ? b >> 8 : b << 8; }

/* ---- attack analysis (attacks.rs:101-200) ------------------------------ */
struct Analysis {
  u64 our, their, occ;
  u64 safe_king;  /* AttackInfo::safe_king_squares */
  u64 checkers;   /* AttackInfo::checkers */
  u64 pins;       /* AttackInfo::pins */
  u64 attacks;    /* AttackInfo::attacks restricted to what move generation reads: exact on the king's
                     neighbour squares and on the castling walks, sliders seen through our king (see below) */
  u64 block;      /* blocking_ray of position.rs:333-353; 0 when doubly checked */
  int us, king;
};

/* Differences from the reference's loop, none observable:
 *  - sliders look through our king (occupancy without it) for every slider,
 *    not only for checking ones: a slider's attack set changes when the king is
 *    removed only if the king is in it, i.e. only for checkers, and for those
 *    the reference subtracts exactly that extended set (attacks.rs:143,163,183).
 *    `attacks` is only used for castling, which requires no checkers, where both
 *    definitions coincide.
 *  - pins are found from the king: our first blockers on its lines, then the
 *    enemy sliders that appear when they are removed; "exactly one blocker and
 *    it is ours" (attacks.rs:147-156) is the same condition.
 *  - xrays (attacks.rs:93) are not computed: move generation never reads them. */
template <int US = -1, class TB>
PB_HD void analyze(const Pos& p, const TB& tb, Analysis& a, bool want_attacks = true) {
  const int us = US >= 0 ? US : meta_stm(p.meta); /* US: side to move when known at compile time */
  const int them = us ^ 1;
  const u64 occ = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  const u64 our = us == 0 ? p.white : occ & ~p.white;
  const u64 their = occ ^ our;
  const u64 king_bb = p.kings & our;
  const int king = lsb(king_bb);
  const u64 occ_nk = occ ^ king_bb;
  const u64 their_diag = (p.bishops | p.queens) & their;
  const u64 their_orth = (p.rooks | p.queens) & their;

  const u64 king_diag = tb.bishop(king, occ);
  const u64 king_orth = tb.rook(king, occ);
  u64 checkers = (tb.knight(king) & p.knights & their) |
                 (pawn_attacks(king_bb, us) & p.pawns & their) | (king_diag & their_diag) |
                 (king_orth & their_orth);

  /* The enemy attack map is only ever read on the squares the king could step to and on the castling
   * walks (generate_king_moves / generate_castle_moves): it is built from the enemy pieces that can reach
   * those squares, found with the per-king-square zone tables (see count_moves_white). */
  const u64 king_targets = tb.king(king) & ~our;
  u64 need = king_targets;
  if (checkers == 0) {
    const int rights = meta_castling(p.meta) >> (2 * us), sh = 56 * us;
    if ((rights & 1) && (occ & (0x60ULL << sh)) == 0) need |= 0x60ULL << sh;
    if ((rights & 2) && (occ & (0x0EULL << sh)) == 0) need |= 0x0CULL << sh;
  }
  u64 att = 0;
  if (need && want_attacks) { /* a caller that will emit neither king nor castling moves skips the map */
    att = tb.king(lsb(p.kings & their)) | pawn_attacks(p.pawns & their, them);
    for (Squares b(p.knights & their & tb.s->knight_zone[king]); b.any();) att |= tb.knight(b.pop());
    for (Squares b(their_diag & tb.s->diag_zone[king]); b.any();) {
      const int sq = b.pop();
      if ((tb.s->diag[sq] | tb.s->anti[sq]) & need) att |= tb.bishop(sq, occ_nk);
    }
    for (Squares b(their_orth & tb.s->orth_zone[king]); b.any();) {
      const int sq = b.pop();
      if (tb.s->orth[sq] & need) att |= tb.rook(sq, occ_nk);
    }
  }

  u64 pins = 0;
  { /* the x-ray lookup only when an enemy slider stands on one of the king's lines behind our piece */
    const u64 cand = king_diag & our;
    if (cand && ((tb.s->diag[king] | tb.s->anti[king]) & their_diag & ~king_diag)) {
      u64 pinners = tb.bishop(king, occ ^ cand) & ~king_diag & their_diag;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), king) & cand;
    }
  }
  {
    const u64 cand = king_orth & our;
    if (cand && (tb.s->orth[king] & their_orth & ~king_orth)) {
      u64 pinners = tb.rook(king, occ ^ cand) & ~king_orth & their_orth;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), king) & cand;
    }
  }

  a.us = us;
  a.king = king;
  a.our = our;
  a.their = their;
  a.occ = occ;
  a.attacks = att;
  a.safe_king = tb.king(king) & ~our & ~att;
  a.checkers = checkers;
  a.pins = pins;
  /* position.rs:333-358 */
  if (checkers == 0) {
    a.block = ~0ULL;
  } else if ((checkers & (checkers - 1)) == 0) {
    u64 r = tb.ray(lsb(checkers), king);
    a.block = r ? r : checkers;
  } else {
    a.block = 0;
  }
}

/* AttackInfo exactly as the reference publishes it (attacks.rs:89-200, Position::attack_info
 * position.rs:285-292): `attacks` with the full occupancy (sliders do NOT look through the king),
 * `xrays` included, `safe_king_squares` with the checking sliders' rays extended through the king.
 * Move generation uses analyze() above; this is the API-level picture the reference's tests pin
 * (attacks.rs:555-695). */
struct AttackInfoOut {
  u64 attacks, checkers, pins, xrays, safe_king_squares;
};
template <class TB>
PB_HD AttackInfoOut attack_info(const Pos& p, const TB& tb) {
  const int us = meta_stm(p.meta), them = us ^ 1;
  const u64 occ = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  const u64 our = us == 0 ? p.white : occ & ~p.white;
  const u64 their = occ ^ our;
  const u64 king_bb = p.kings & our;
  const int king = lsb(king_bb);
  const u64 occ_nk = occ ^ king_bb;
  AttackInfoOut r;
  r.attacks = tb.king(lsb(p.kings & their)) | pawn_attacks(p.pawns & their, them);
  r.checkers = pawn_attacks(king_bb, us) & p.pawns & their;
  r.pins = r.xrays = 0;
  u64 safe = tb.king(king) & ~our;
  for (u64 b = p.knights & their; b; b &= b - 1) {
    const u64 t = tb.knight(lsb(b));
    r.attacks |= t;
    if (t & king_bb) r.checkers |= b & (0 - b);
  }
  /* queens, bishops, rooks (attacks.rs:138-197); the blocker test uses the ray towards the king */
  for (int diag = 0; diag < 2; diag++) {
    for (u64 b = (diag ? p.bishops | p.queens : p.rooks | p.queens) & their; b; b &= b - 1) {
      const int sq = lsb(b);
      const u64 t = diag ? tb.bishop(sq, occ) : tb.rook(sq, occ);
      r.attacks |= t;
      if (t & king_bb) {
        r.checkers |= b & (0 - b);
        safe &= ~(diag ? tb.bishop(sq, occ_nk) : tb.rook(sq, occ_nk));
        continue;
      }
      /* bishop_ray / rook_ray (attacks.rs:58-64): the ray only when aligned the slider's way; a queen takes part in
       * both passes, and is aligned with the king in at most one of them */
      const bool same_line = ((sq ^ king) & 7) == 0 || ((sq ^ king) & 56) == 0;
      const u64 ray = (diag ? !same_line : same_line) ? tb.ray(sq, king) : 0;
      const u64 blockers = (ray & occ) & ~(b & (0 - b));
      if (blockers && (blockers & (blockers - 1)) == 0) {
        if (blockers & our) r.pins |= blockers;
        else r.xrays |= blockers;
      }
    }
  }
  r.safe_king_squares = safe & ~r.attacks;
  return r;
}

/* position.rs:1226-1288 with the walk masks of attacks.rs:203-214.  Returns a
 * 2-bit set: bit 0 short, bit 1 long. */
PB_HD int castle_moves(const Pos& p, const Analysis& a) {
  if (a.checkers) return 0;
  const int rights = meta_castling(p.meta) >> (2 * a.us);
  const int sh = 56 * a.us;
  int r = 0;
  if ((rights & 1) && ((a.attacks | a.occ) & (0x60ULL << sh)) == 0) r |= 1;
  if ((rights & 2) && (a.attacks & (0x0CULL << sh)) == 0 && (a.occ & (0x0EULL << sh)) == 0) r |= 2;
  return r;
}

/* En passant (position.rs:1144-1178).  Returns the set of our pawns that may
 * capture on the en passant square. */
template <class TB>
PB_HD u64 en_passant_sources(const Pos& p, const TB& tb, const Analysis& a) {
  const int ep = meta_ep(p.meta);
  if (ep >= NO_SQUARE) return 0;
  const int them = a.us ^ 1;
  const u64 ep_bb = bit(ep);
  const u64 candidates = pawn_attacks(ep_bb, them) & p.pawns & a.our;
  if (!candidates) return 0;
  const u64 ep_pawn = push(ep_bb, them); /* the pushed pawn */
  if (a.checkers & ep_pawn) return candidates & ~a.pins;
  const u64 their_diag = (p.bishops | p.queens) & a.their;
  const u64 their_orth = (p.rooks | p.queens) & a.their;
  u64 ok = 0;
  for (u64 b = candidates; b; b &= b - 1) {
    const u64 cand = b & (0 - b);
    const u64 after = ((a.occ & ~cand) & ~ep_pawn) | ep_bb;
    if ((tb.bishop(a.king, after) & their_diag) == 0 && (tb.rook(a.king, after) & their_orth) == 0)
      ok |= cand;
  }
  return ok;
}

/* Pawn source sets after the pin filter (position.rs:1128,1200,1217 applied
 * set-wise): a pinned pawn may push only along the king's file and capture only
 * along the pin diagonal. */
struct PawnSets {
  u64 west, east, pushers;
};
template <class TB>
PB_HD PawnSets pawn_sources(const Pos& p, const TB& tb, const Analysis& a) {
  const u64 pawns = p.pawns & a.our;
  const u64 free_pawns = pawns & ~a.pins;
  const u64 pinned = pawns & a.pins;
  PawnSets s;
  s.west = s.east = s.pushers = free_pawns;
  if (pinned) {
    const u64 d = tb.s->diag[a.king], an = tb.s->anti[a.king];
    s.pushers |= pinned & (FILE_A << (a.king & 7));
    /* white: west capture (+7) runs along the anti-diagonal, east (+9) along the
     * diagonal; black: west (-9) along the diagonal, east (-7) along the anti-diagonal */
    s.west |= pinned & (a.us == 0 ? an : d);
    s.east |= pinned & (a.us == 0 ? d : an);
  }
  return s;
}

/* ---- count only: == generate_moves().len() (position.rs:914-916) --------
 * This is the function perft spends its time in (one call per node of the last
 * ply but one), so it is organised to execute as few instructions as possible:
 *  - the position is first mirrored so that the side to move plays up the board
 *    (byte-swapped bitboards, swapped castling nibble): a count does not depend
 *    on the colour, and all colour-dependent shifts and masks become constants;
 *  - the enemy attack map is only needed on the squares the king could step to
 *    and on the castling walk; enemy sliders whose empty-board lines miss those
 *    squares are skipped without a table lookup, and when the king has no
 *    candidate square at all the whole map is skipped;
 *  - the x-ray lookups for pins run only when an enemy slider stands on one of
 *    the king's empty-board lines behind one of our pieces;
 *  - our sliders hemmed in by our own pieces are skipped.
 * Every shortcut is exact; tests/test_device_logic_on_host.py compares the
 * result with the oracle's list length on every node of the reference trees. */
/* the squares with a rank/file (diagonal) neighbour in `set`: a slider none of whose neighbours is free of
 * its own pieces has no moves, and its attack set holds own pieces only */
PB_HD u64 orth_neighbours(u64 set) { return (set >> 8) | (set << 8) | ((set >> 1) & ~FILE_H) | ((set << 1) & ~FILE_A); }
PB_HD u64 diag_neighbours(u64 set) { return (((set >> 9) | (set << 7)) & ~FILE_H) | (((set >> 7) | (set << 9)) & ~FILE_A); }
PB_HD u64 bswap64(u64 x) {
#ifdef __CUDA_ARCH__
  const u32 lo = (u32)x, hi = (u32)(x >> 32);
  return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
  return __builtin_bswap64(x);
#endif
}

/* The count for a position that is already "white to move": `our` = the mover's occupancy, the
 * piece sets merged over both colours, `rights` = the mover's castling bits (1 short, 2 long),
 * `ep` = en passant square (>= 64: none). */
template <class TB>
PB_HD u32 count_moves_white(u64 our, u64 pawns, u64 knights, u64 bishops, u64 rooks, u64 queens, u64 kings, int rights,
                            int ep, const TB& tb, int known_king = -1, int known_their_king = -1) {
  const u64 occ = pawns | knights | bishops | rooks | queens | kings;
  const u64 their = occ ^ our;
  const u64 king_bb = kings & our;
  const int king = known_king >= 0 ? known_king : lsb(king_bb);
  const u64 their_diag = (bishops | queens) & their;
  const u64 their_orth = (rooks | queens) & their;

  const u64 king_diag = tb.bishop(king, occ);
  const u64 king_orth = tb.rook(king, occ);
  const u64 checkers = (tb.knight(king) & knights & their) | (pawn_attacks(king_bb, 0) & pawns & their) |
                       (king_diag & their_diag) | (king_orth & their_orth);

  /* king steps and castling: the only consumers of the enemy attack map */
  const u64 king_targets = tb.king(king) & ~our;
  /* castling walks that are worth checking: right held, not in check, squares between king and rook empty
   * (position.rs:1226-1288, walk masks attacks.rs:203-214) */
  u64 castle_walks = 0;
  if (rights && checkers == 0) {
    if ((rights & 1) && (occ & 0x60ULL) == 0) castle_walks = 0x60ULL;
    if ((rights & 2) && (occ & 0x0EULL) == 0) castle_walks |= 0x0CULL;
  }
  const u64 need = king_targets | castle_walks;
  u32 n = 0;
  if (need) {
    /* `rem` = squares of `need` no enemy piece attacks; it only shrinks, so every later piece is
     * tested against fewer squares and the scan stops when nothing is left */
    const u64 occ_nk = occ ^ king_bb;
    const int their_king = known_their_king >= 0 ? known_their_king : lsb(kings & their);
    u64 rem = need & ~(tb.king(their_king) | pawn_attacks(pawns & their, 1));
    for (Squares b(knights & their & tb.s->knight_zone[king]); b.any();) rem &= ~tb.knight(b.pop());
    for (Squares b(their_diag & tb.s->diag_zone[king]); b.any() && rem;) {
      const int sq = b.pop();
      if ((tb.s->diag[sq] | tb.s->anti[sq]) & rem) rem &= ~tb.bishop(sq, occ_nk);
    }
    for (Squares b(their_orth & tb.s->orth_zone[king]); b.any() && rem;) {
      const int sq = b.pop();
      if (tb.s->orth[sq] & rem) rem &= ~tb.rook(sq, occ_nk);
    }
    n = popc(king_targets & rem);
    if (castle_walks) { /* both squares of a walk must be unattacked */
      const u64 safe_walks = castle_walks & rem;
      n += ((safe_walks & 0x60ULL) == 0x60ULL) + ((safe_walks & 0x0CULL) == 0x0CULL);
    }
  }

  u64 block = ~0ULL;
  if (checkers) {
    if (checkers & (checkers - 1)) return n; /* double check: king moves only, position.rs:356 */
    const u64 r = tb.ray(lsb(checkers), king);
    block = r ? r : checkers;
  }

  u64 pins = 0;
  {
    const u64 cand = king_diag & our;
    if (cand && ((tb.s->diag[king] | tb.s->anti[king]) & their_diag & ~king_diag)) {
      u64 pinners = tb.bishop(king, occ ^ cand) & ~king_diag & their_diag;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), king) & cand;
    }
  }
  {
    const u64 cand = king_orth & our;
    if (cand && (tb.s->orth[king] & their_orth & ~king_orth)) {
      u64 pinners = tb.rook(king, occ ^ cand) & ~king_orth & their_orth;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), king) & cand;
    }
  }

  const u64 not_our = ~our;
  const u64 targets = not_our & block;
  for (Squares b(knights & our & ~pins); b.any();) n += popc(tb.knight(b.pop()) & targets);
  /* sliders hemmed in by our own pieces (no neighbour on their lines that is not ours) have no
   * moves and need no lookup; found set-wise */
  const u64 orth_open = orth_neighbours(not_our), diag_open = diag_neighbours(not_our);
  for (Squares b((rooks | queens) & our & orth_open); b.any();) {
    const int from = b.pop();
    u64 t = tb.rook(from, occ) & targets;
    if ((pins >> from) & 1) t &= tb.line(king, from);
    n += popc(t);
  }
  for (Squares b((bishops | queens) & our & diag_open); b.any();) {
    const int from = b.pop();
    u64 t = tb.bishop(from, occ) & targets;
    if ((pins >> from) & 1) t &= tb.line(king, from);
    n += popc(t);
  }

  /* pawns, set-wise (position.rs:1106-1224); a pinned pawn pushes only along the
   * king's file and captures only along the pin diagonal */
  const u64 our_pawns = pawns & our;
  u64 west_src = our_pawns & ~pins, east_src = west_src, pushers = west_src;
  const u64 pinned = our_pawns & pins;
  if (pinned) {
    pushers |= pinned & (FILE_A << (king & 7));
    west_src |= pinned & tb.s->anti[king];
    east_src |= pinned & tb.s->diag[king];
  }
  const u64 capture_targets = their & block;
  const u64 west = ((west_src & ~FILE_A) << 7) & capture_targets;
  const u64 east = ((east_src & ~FILE_H) << 9) & capture_targets;
  const u64 single_all = (pushers << 8) & ~occ;
  const u64 single = single_all & block;
  const u64 dbl = ((single_all & RANK_3) << 8) & ~occ & block;
  n += popc(west) + popc(east) + popc(single) + popc(dbl);
  if ((west | east | single) & RANK_8) n += 3 * (popc(west & RANK_8) + popc(east & RANK_8) + popc(single & RANK_8));

  if (ep < NO_SQUARE) { /* en passant, position.rs:1144-1178 */
    const u64 ep_bb = bit(ep);
    const u64 candidates = pawn_attacks(ep_bb, 1) & our_pawns;
    if (candidates) {
      const u64 ep_pawn = ep_bb >> 8;
      if (checkers & ep_pawn) {
        n += popc(candidates & ~pins);
      } else {
        for (u64 b = candidates; b; b &= b - 1) {
          const u64 after = ((occ & ~(b & (0 - b))) & ~ep_pawn) | ep_bb;
          n += (tb.bishop(king, after) & their_diag) == 0 && (tb.rook(king, after) & their_orth) == 0;
        }
      }
    }
  }
  return n;
}

/* The board seen from the other side: ranks mirrored, colours swapped, castling nibbles swapped,
 * en passant square mirrored.  Legal move counts are invariant under it. */
PB_HD Pos mirror(const Pos& p) {
  Pos r;
  const u64 all = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  r.white = bswap64(all & ~p.white);
  r.pawns = bswap64(p.pawns), r.knights = bswap64(p.knights), r.bishops = bswap64(p.bishops);
  r.rooks = bswap64(p.rooks), r.queens = bswap64(p.queens), r.kings = bswap64(p.kings);
  const int c = meta_castling(p.meta), ep = meta_ep(p.meta);
  r.meta = make_meta(((c >> 2) | (c << 2)) & 15, ep < NO_SQUARE ? ep ^ 56 : NO_SQUARE, meta_stm(p.meta) ^ 1,
                     meta_halfmove(p.meta), meta_fullmove(p.meta));
  return r;
}

/* K2 keeps both king squares in spare bits of the meta word of the parent it holds in shared memory
 * (bits 48-53: the mover's king, 54-59: the other king), so that the per-child count does not have to
 * find them again: make_move_kind<US, true> maintains them. */
constexpr int META_MOVER_KING = 48, META_OTHER_KING = 54;
PB_HD u64 with_king_squares(const Pos& p) {
  const u64 occ = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  const u64 mover = meta_stm(p.meta) == 0 ? p.white : occ & ~p.white;
  return (p.meta & 0x0000FFFFFFFFFFFFULL) | ((u64)lsb(p.kings & mover) << META_MOVER_KING) |
         ((u64)lsb(p.kings & ~mover) << META_OTHER_KING);
}

/* The count for a child produced by make_move_kind<1, true> from a black-to-move parent that carries
 * its king squares: white to move, our king = the parent's other king. */
template <class TB>
PB_HD u32 count_child_white(const Pos& q, const TB& tb) {
  return count_moves_white(q.white, q.pawns, q.knights, q.bishops, q.rooks, q.queens, q.kings, meta_castling(q.meta) & 3,
                           meta_ep(q.meta), tb, (int)(q.meta >> META_OTHER_KING) & 63, (int)(q.meta >> META_MOVER_KING) & 63);
}

/* generate_moves().len() of any position */
template <class TB>
PB_HD u32 count_moves(const Pos& p, const TB& tb) {
  if (meta_stm(p.meta) == 0)
    return count_moves_white(p.white, p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings, meta_castling(p.meta) & 3,
                             meta_ep(p.meta), tb);
  const Pos m = mirror(p);
  return count_moves_white(m.white, m.pawns, m.knights, m.bishops, m.rooks, m.queens, m.kings, meta_castling(m.meta) & 3,
                           meta_ep(m.meta), tb);
}

/* ---- ordered generation (position.rs:317-408; order of SURVEY.md A.3) ---
 * emit(from, to, promotion, kind) is called once per legal move in the reference's
 * order; kind = KIND_* of the moving piece.  Returns the number of moves. */
template <int US = -1, class TB, class Emit>
PB_HD u32 generate_moves(const Pos& p, const TB& tb, Emit&& emit) {
  Analysis a;
  analyze<US>(p, tb, a);
  u32 n = 0;
  /* generate_king_moves, position.rs:1034-1040 */
  for (Squares t(a.safe_king); t.any(); ++n) emit(a.king, t.pop(), 0, KIND_KING);
  if (a.block == 0) return n;
  const u64 targets = ~a.our & a.block;

  /* generate_knight_moves, position.rs:1042-1059 */
  for (Squares b(p.knights & a.our & ~a.pins); b.any();) {
    const int from = b.pop();
    for (Squares t(tb.knight(from) & targets); t.any(); ++n) emit(from, t.pop(), 0, KIND_KNIGHT);
  }
  /* generate_rook_moves over rooks|queens, position.rs:1061-1081 */
  for (Squares b((p.rooks | p.queens) & a.our); b.any();) {
    const int from = b.pop();
    u64 t = tb.rook(from, a.occ) & targets;
    if ((a.pins >> from) & 1) t &= tb.line(a.king, from);
    const int kind = (p.queens >> from) & 1 ? KIND_QUEEN : KIND_ROOK;
    for (Squares ts(t); ts.any(); ++n) emit(from, ts.pop(), 0, kind);
  }
  /* generate_bishop_moves over bishops|queens, position.rs:1083-1104 */
  for (Squares b((p.bishops | p.queens) & a.our); b.any();) {
    const int from = b.pop();
    u64 t = tb.bishop(from, a.occ) & targets;
    if ((a.pins >> from) & 1) t &= tb.line(a.king, from);
    const int kind = (p.queens >> from) & 1 ? KIND_QUEEN : KIND_BISHOP;
    for (Squares ts(t); ts.any(); ++n) emit(from, ts.pop(), 0, kind);
  }

  /* generate_pawn_moves, position.rs:1106-1224 */
  const PawnSets ps = pawn_sources(p, tb, a);
  const int last_rank = a.us == 0 ? 7 : 0;
  const u64 capture_targets = a.their & a.block;
  /* captures: per pawn ascending, per target ascending (:1123-1142) */
  for (Squares b(ps.west | ps.east); b.any();) {
    const int from = b.pop();
    const u64 from_bb = bit(from);
    const u64 t = (pawn_attacks_west(from_bb & ps.west, a.us) | pawn_attacks_east(from_bb & ps.east, a.us)) & capture_targets;
    for (Squares ts(t); ts.any();) {
      const int to = ts.pop();
      if ((to >> 3) == last_rank) {
        emit(from, to, PROMO_Q, KIND_PAWN), emit(from, to, PROMO_R, KIND_PAWN), emit(from, to, PROMO_B, KIND_PAWN),
            emit(from, to, PROMO_N, KIND_PAWN);
        n += 4;
      } else {
        emit(from, to, 0, KIND_PAWN), ++n;
      }
    }
  }
  /* en passant (:1144-1178) */
  if (meta_ep(p.meta) < NO_SQUARE)
    for (u64 b = en_passant_sources(p, tb, a); b; b &= b - 1, ++n) emit(lsb(b), meta_ep(p.meta), 0, KIND_PAWN);
  /* single pushes in ascending target order (:1180-1204) */
  const u64 empty = ~a.occ;
  const u64 single_all = push(ps.pushers, a.us) & empty;
  const int back = a.us == 0 ? -8 : 8;
  for (Squares t(single_all & a.block); t.any();) {
    const int to = t.pop();
    if ((to >> 3) == last_rank) {
      emit(to + back, to, PROMO_Q, KIND_PAWN), emit(to + back, to, PROMO_R, KIND_PAWN),
          emit(to + back, to, PROMO_B, KIND_PAWN), emit(to + back, to, PROMO_N, KIND_PAWN);
      n += 4;
    } else {
      emit(to + back, to, 0, KIND_PAWN), ++n;
    }
  }
  /* double pushes in ascending target order (:1205-1223) */
  for (Squares t(push(single_all & (a.us == 0 ? RANK_3 : RANK_6), a.us) & empty & a.block); t.any(); ++n) {
    const int to = t.pop();
    emit(to + 2 * back, to, 0, KIND_PAWN);
  }
  /* generate_castle_moves: short, then long (:1236-1286) */
  const int c = castle_moves(p, a);
  const int home = 4 + 56 * a.us;
  if (c & 1) emit(home, home + 2, 0, KIND_KING), ++n;
  if (c & 2) emit(home, home - 2, 0, KIND_KING), ++n;
  return n;
}

/* ---- make_move (position.rs:414-433 and helpers :435-732) --------------- */
template <class TB>
PB_HD void make_move(Pos& p, const TB& tb, u16 mv) {
  const int from = mv & 63, to = (mv >> 6) & 63, promo = (mv >> 12) & 7;
  const u64 fb = bit(from), tbit = bit(to);
  const int us = meta_stm(p.meta);
  int castling = meta_castling(p.meta);
  const int prev_ep = meta_ep(p.meta);
  int halfmove = (meta_halfmove(p.meta) + 1) & 255; /* :419, u8 */
  int fullmove = meta_fullmove(p.meta);
  int ep = NO_SQUARE;

  /* update_castling_rights, :435-468 */
  castling &= tb.s->keep_from[from] & tb.s->keep_to[to];

  const u64 occ = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  const u64 our = us == 0 ? p.white : occ & ~p.white;
  u64 white = p.white;

  /* handle_capture, :470-502 (a king is never captured) */
  if ((occ & ~our) & tbit) {
    halfmove = 0;
    p.pawns &= ~tbit, p.knights &= ~tbit, p.bishops &= ~tbit, p.rooks &= ~tbit, p.queens &= ~tbit;
    white &= ~tbit;
  }

  if (p.pawns & fb) { /* make_pawn_move, :504-618 */
    halfmove = 0;
    if (to == prev_ep) { /* prev_ep == 64 never equals a square */
      const u64 captured = bit((to & 7) + (from & 56));
      p.pawns &= ~captured;
      white &= ~captured;
    }
    p.pawns ^= fb;
    if (promo) {
      if (promo == PROMO_Q) p.queens |= tbit;
      else if (promo == PROMO_R) p.rooks |= tbit;
      else if (promo == PROMO_B) p.bishops |= tbit;
      else p.knights |= tbit;
    } else {
      p.pawns |= tbit;
      /* :606-615: the en passant square is recorded only when an enemy pawn attacks it */
      if ((from ^ to) == 16) { /* the pawns that attack the skipped square stand beside the pushed pawn */
        const u64 beside = ((tbit & ~FILE_H) << 1) | ((tbit & ~FILE_A) >> 1);
        if (beside & p.pawns & ~our) ep = (from + to) >> 1;
      }
    }
  } else if (p.kings & fb) { /* make_king_move, :622-698 */
    const int backrank = 56 * us;
    if (from == backrank + 4 && (to & 56) == backrank) {
      u64 rook_move = 0;
      if ((to & 7) == 6) rook_move = (0xA0ULL) << backrank;      /* h -> f */
      else if ((to & 7) == 2) rook_move = (0x09ULL) << backrank; /* a -> d */
      /* the reference clears OUR rook on the home square (if there is one: the
       * castling right is trusted, :632-678) and puts a rook on the target square */
      if (rook_move) {
        const u64 home = rook_move & (0x81ULL << backrank), dest = rook_move ^ home;
        const u64 our_home_rook = p.rooks & our & home;
        p.rooks = (p.rooks ^ our_home_rook) | dest;
        if (us == 0) white = (white ^ our_home_rook) | dest;
      }
    }
    p.kings ^= fb | tbit;
  } else { /* make_regular_move, :700-732 */
    const u64 m = fb | tbit;
    if (p.queens & fb) p.queens ^= m;
    else if (p.rooks & fb) p.rooks ^= m;
    else if (p.bishops & fb) p.bishops ^= m;
    else if (p.knights & fb) p.knights ^= m;
  }
  if (us == 0) white = (white & ~fb) | tbit;
  p.white = white;
  if (us == 1) fullmove = (fullmove + 1) & 0xFFFF; /* :428-430 */
  p.meta = make_meta(castling, ep, us ^ 1, halfmove, fullmove);
}

/* make_move with the moving piece's kind known (it comes with every generated move).  The counting
 * path does not maintain the clocks (halfmove, fullmove: position.rs:419,428-430) because a count does
 * not read them; K1 (CLOCKS) does, its children are complete records.  Board, castling rights, en passant square and side to move are exactly those of
 * make_move (checked node by node in tests/test_device_logic_on_host.py). */
template <int US = -1, bool TRACK_KINGS = false, bool CLOCKS = false, class TB>
PB_HD void make_move_kind(Pos& p, const TB& tb, u16 mv, int kind) {
  const int from = mv & 63, to = (mv >> 6) & 63, promo = (mv >> 12) & 7;
  const u64 fb = bit(from), tbit = bit(to), m = fb | tbit;
  const int us = US >= 0 ? US : meta_stm(p.meta); /* US: the mover's colour when known at compile time */
  const int castling = meta_castling(p.meta) & tb.s->keep_from[from] & tb.s->keep_to[to];
  const int prev_ep = meta_ep(p.meta);
  int ep = NO_SQUARE;
  u64 white = p.white & ~tbit;
  /* CLOCKS (K1, whose children are complete records): position.rs:419 and the resets of :470-502 / :504-618 */
  int halfmove = 0, fullmove = 0;
  if (CLOCKS) {
    const bool resets = kind == KIND_PAWN || ((p.pawns | p.knights | p.bishops | p.rooks | p.queens) & tbit) != 0;
    halfmove = resets ? 0 : (meta_halfmove(p.meta) + 1) & 255;
    fullmove = (meta_fullmove(p.meta) + us) & 0xFFFF; /* :428-430 */
  }
  /* handle_capture: whatever stood on `to` (never a king, never ours) is gone */
  p.pawns &= ~tbit, p.knights &= ~tbit, p.bishops &= ~tbit, p.rooks &= ~tbit, p.queens &= ~tbit;
  if (kind == KIND_PAWN) {
    if (to == prev_ep) {
      const u64 captured = bit((to & 7) + (from & 56));
      p.pawns &= ~captured;
      white &= ~captured;
    }
    p.pawns ^= fb;
    if (promo) {
      if (promo == PROMO_Q) p.queens |= tbit;
      else if (promo == PROMO_R) p.rooks |= tbit;
      else if (promo == PROMO_B) p.bishops |= tbit;
      else p.knights |= tbit;
    } else {
      p.pawns |= tbit;
      if ((from ^ to) == 16) { /* the pawns that attack the skipped square stand beside the pushed pawn */
        const u64 beside = ((tbit & ~FILE_H) << 1) | ((tbit & ~FILE_A) >> 1);
        if (beside & p.pawns & (us == 0 ? ~white : white)) ep = (from + to) >> 1;
      }
    }
  } else if (kind == KIND_KING) {
    const int backrank = 56 * us;
    if (from == backrank + 4 && (to & 56) == backrank && ((to & 7) == 6 || (to & 7) == 2)) {
      const u64 home = ((to & 7) == 6 ? 0x80ULL : 0x01ULL) << backrank, dest = ((to & 7) == 6 ? 0x20ULL : 0x08ULL) << backrank;
      const u64 our_home_rook = p.rooks & home & (us == 0 ? white : ~white);
      p.rooks = (p.rooks ^ our_home_rook) | dest;
      if (us == 0) white = (white ^ our_home_rook) | dest;
    }
    p.kings ^= m;
  } else if (kind == KIND_KNIGHT) {
    p.knights ^= m;
  } else if (kind == KIND_BISHOP) {
    p.bishops ^= m;
  } else if (kind == KIND_ROOK) {
    p.rooks ^= m;
  } else {
    p.queens ^= m;
  }
  if (us == 0) white = (white & ~fb) | tbit;
  p.white = white;
  u64 kings_known = 0;
  if (TRACK_KINGS) { /* bits 48-59 of the parent's meta, the mover's king updated when it moved */
    kings_known = p.meta & (0xFFFULL << META_MOVER_KING);
    if (kind == KIND_KING) kings_known = (kings_known & ~(63ULL << META_MOVER_KING)) | ((u64)to << META_MOVER_KING);
  }
  p.meta = make_meta(castling, ep, us ^ 1, halfmove, fullmove) | kings_known;
}

/* ---- Position::hash (position.rs:110) ------------------------------------------------------ */
/* compute_hash, position.rs:776-806 */
template <class TB>
PB_HD u64 compute_hash(const Pos& p, const TB& tb) {
  const u64* z = tb.g.zobrist;
  u64 key = 0;
  if (meta_stm(p.meta)) key ^= ldg64(z + ZOBRIST_BLACK_TO_MOVE);
  const int castling = meta_castling(p.meta);
  for (int i = 0; i < 4; i++)
    if (castling & (1 << i)) key ^= ldg64(z + ZOBRIST_CASTLE + i);
  if (meta_ep(p.meta) < NO_SQUARE) key ^= ldg64(z + ZOBRIST_EP + (meta_ep(p.meta) & 7));
  const u64 sets[6] = {p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings};
  for (int kind = 0; kind < 6; kind++)
    for (u64 b = sets[kind]; b; b &= b - 1) {
      const int sq = lsb(b);
      key ^= ldg64(z + ((p.white >> sq) & 1 ? 0 : 384) + kind * 64 + sq);
    }
  return key;
}

/* What make_move XORs into the hash for move `mv` played in `p` (the position BEFORE the move):
 * update_castling_rights :442-466, handle_capture :491-497, make_pawn_move :524-531, :536-543, :551-594, :614,
 * make_king_move :640-696, make_regular_move :714-727.  Faithful to what the reference does NOT do as well: the
 * BLACK_TO_MOVE key is never toggled and the key of a previous en passant square is never removed. */
template <class TB>
PB_HD u64 hash_delta(const Pos& p, const TB& tb, u16 mv) {
  const u64* z = tb.g.zobrist;
  const int from = mv & 63, to = (mv >> 6) & 63, promo = (mv >> 12) & 7;
  const u64 fb = bit(from), tbit = bit(to);
  const int us = meta_stm(p.meta);
  const int ours = us * 384, theirs = 384 - ours;
  const u64 occ = p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings;
  const u64 our = us == 0 ? p.white : occ & ~p.white;
  u64 d = 0;
  const int lost = meta_castling(p.meta) & ~(tb.s->keep_from[from] & tb.s->keep_to[to]);
  for (int i = 0; i < 4; i++)
    if (lost & (1 << i)) d ^= ldg64(z + ZOBRIST_CASTLE + i);
  if ((occ & ~our) & tbit) { /* the captured piece: queens, rooks, bishops, knights, pawns (a king is never matched) */
    const int kind = p.queens & tbit ? KIND_QUEEN : p.rooks & tbit ? KIND_ROOK : p.bishops & tbit ? KIND_BISHOP
                   : p.knights & tbit ? KIND_KNIGHT : p.pawns & tbit ? KIND_PAWN : -1;
    if (kind >= 0) d ^= ldg64(z + theirs + kind * 64 + to);
  }
  if (p.pawns & our & fb) {
    if (to == meta_ep(p.meta)) d ^= ldg64(z + theirs + KIND_PAWN * 64 + (to & 7) + (from & 56));
    d ^= ldg64(z + ours + KIND_PAWN * 64 + from);
    if (promo) { /* Promotion and PieceKind share their numbering: 1 knight .. 4 queen */
      d ^= ldg64(z + ours + promo * 64 + to);
    } else {
      d ^= ldg64(z + ours + KIND_PAWN * 64 + to);
      if ((from ^ to) == 16) {
        const u64 beside = ((tbit & ~FILE_H) << 1) | ((tbit & ~FILE_A) >> 1);
        if (beside & p.pawns & ~our & ~tbit) d ^= ldg64(z + ZOBRIST_EP + (to & 7));
      }
    }
  } else if (p.kings & our & fb) {
    const int backrank = 56 * us;
    if (from == backrank + 4 && (to & 56) == backrank) { /* the rook keys are toggled whether or not a rook stands there */
      if ((to & 7) == 6) d ^= ldg64(z + ours + KIND_ROOK * 64 + backrank + 7) ^ ldg64(z + ours + KIND_ROOK * 64 + backrank + 5);
      else if ((to & 7) == 2) d ^= ldg64(z + ours + KIND_ROOK * 64 + backrank) ^ ldg64(z + ours + KIND_ROOK * 64 + backrank + 3);
    }
    d ^= ldg64(z + ours + KIND_KING * 64 + from) ^ ldg64(z + ours + KIND_KING * 64 + to);
  } else {
    const int kind = p.queens & our & fb ? KIND_QUEEN : p.rooks & our & fb ? KIND_ROOK : p.bishops & our & fb ? KIND_BISHOP
                   : p.knights & our & fb ? KIND_KNIGHT : -1;
    if (kind >= 0) d ^= ldg64(z + ours + kind * 64 + from) ^ ldg64(z + ours + kind * 64 + to);
  }
  return d;
}

} /* namespace pabi */

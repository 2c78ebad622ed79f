/*
 * tables.c -- regenerates pabi's precomputed bitboard tables from their
 * mathematical definition (TEST INFRASTRUCTURE; see pabi_oracle.h).
 *
 * Follows /root/reference/src/chess/generated.rs:30-82 (what tables exist, their
 * sizes and index schemes) and the semantics pinned by attacks.rs:224-553.
 * generated/rook_attacks.rs is missing from the mounted reference
 * (.MISSING_LARGE_BLOBS); it is rebuilt here with the same recipe that reproduces
 * the present generated/bishop_attacks.rs bit for bit (checked in
 * tests/test_oracle_tables.py against tests/golden/reference_tables.npz).
 */
#include "oracle_internal.h"

#include <string.h>

uint64_t ORC_KNIGHT_ATTACKS[64], ORC_KING_ATTACKS[64];
uint64_t ORC_WHITE_PAWN_ATTACKS[64], ORC_BLACK_PAWN_ATTACKS[64];
uint64_t ORC_BISHOP_MASKS[64], ORC_ROOK_MASKS[64];
uint64_t ORC_BISHOP_OFFSETS[64], ORC_ROOK_OFFSETS[64];
uint64_t ORC_BISHOP_ATTACKS[ORC_BISHOP_ATTACKS_COUNT];
uint64_t ORC_ROOK_ATTACKS[ORC_ROOK_ATTACKS_COUNT];
uint64_t ORC_RAYS[4096], ORC_BISHOP_RAYS[4096], ORC_ROOK_RAYS[4096];
uint64_t ORC_PIECE_KEYS[768], ORC_EP_KEYS[8], ORC_MISC_KEYS[5];

static const int BISHOP_DIRS[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
static const int ROOK_DIRS[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};

static int on_board(int f, int r) { return f >= 0 && f < 8 && r >= 0 && r < 8; }

/* Attack set of a slider on `sq` for occupancy `occ`: walk each direction and stop
 * on (and include) the first occupied square (Appendix A.2 of SURVEY.md;
 * pictures in attacks.rs:262-301). */
static uint64_t slide(int sq, uint64_t occ, const int dirs[4][2]) {
  uint64_t out = 0;
  for (int d = 0; d < 4; d++) {
    int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
    while (on_board(f, r)) {
      uint64_t bit = 1ULL << (r * 8 + f);
      out |= bit;
      if (occ & bit) break;
      f += dirs[d][0];
      r += dirs[d][1];
    }
  }
  return out;
}

/* Relevant occupancy: ray squares without the last (edge) square of each ray
 * (attacks.rs:249-288 pictures for E4). */
static uint64_t relevant(int sq, const int dirs[4][2]) {
  uint64_t out = 0;
  for (int d = 0; d < 4; d++) {
    int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
    while (on_board(f, r) && on_board(f + dirs[d][0], r + dirs[d][1])) {
      out |= 1ULL << (r * 8 + f);
      f += dirs[d][0];
      r += dirs[d][1];
    }
  }
  return out;
}

/* Inverse of pext (attacks.rs:75-86 packs masked bits LS1B first). */
static uint64_t deposit(uint64_t idx, uint64_t mask) {
  uint64_t out = 0;
  while (mask) {
    uint64_t ls1b = mask & (~mask + 1);
    if (idx & 1) out |= ls1b;
    idx >>= 1;
    mask ^= ls1b;
  }
  return out;
}

static void build_slider(const int dirs[4][2], uint64_t *masks, uint64_t *offsets,
                         uint64_t *table, size_t expect) {
  size_t off = 0;
  for (int sq = 0; sq < 64; sq++) {
    masks[sq] = relevant(sq, dirs);
    offsets[sq] = off;
    size_t n = (size_t)1 << orc_popcount(masks[sq]);
    for (size_t i = 0; i < n; i++) table[off + i] = slide(sq, deposit(i, masks[sq]), dirs);
    off += n;
  }
  if (off != expect) __builtin_trap();
}

/* RAYS[from*64+to]: squares from `from` (inclusive) towards `to` (exclusive) when
 * both lie on one rank, file or diagonal; empty otherwise and for from == to
 * (attacks.rs:486-553). */
static void build_rays(void) {
  for (int from = 0; from < 64; from++)
    for (int to = 0; to < 64; to++) {
      int df = (to & 7) - (from & 7), dr = (to >> 3) - (from >> 3);
      uint64_t ray = 0;
      int diag = (df != 0 && (df == dr || df == -dr));
      int orth = ((df == 0) != (dr == 0));
      if (diag || orth) {
        int sf = (df > 0) - (df < 0), sr = (dr > 0) - (dr < 0);
        int f = from & 7, r = from >> 3;
        while (r * 8 + f != to) {
          ray |= 1ULL << (r * 8 + f);
          f += sf;
          r += sr;
        }
      }
      ORC_RAYS[from * 64 + to] = ray;
      ORC_BISHOP_RAYS[from * 64 + to] = diag ? ray : 0;
      ORC_ROOK_RAYS[from * 64 + to] = orth ? ray : 0;
    }
}

static void build_leapers(void) {
  static const int KN[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
  for (int sq = 0; sq < 64; sq++) {
    int f = sq & 7, r = sq >> 3;
    uint64_t kn = 0, kg = 0, wp = 0, bp = 0;
    for (int i = 0; i < 8; i++)
      if (on_board(f + KN[i][0], r + KN[i][1])) kn |= 1ULL << ((r + KN[i][1]) * 8 + f + KN[i][0]);
    for (int df = -1; df <= 1; df++)
      for (int dr = -1; dr <= 1; dr++)
        if ((df || dr) && on_board(f + df, r + dr)) kg |= 1ULL << ((r + dr) * 8 + f + df);
    /* Pawn tables are all-zero on ranks 1 and 8 for both colours
     * (attacks.rs:411-417). */
    if (r >= 1 && r <= 6) {
      for (int df = -1; df <= 1; df += 2) {
        if (on_board(f + df, r + 1)) wp |= 1ULL << ((r + 1) * 8 + f + df);
        if (on_board(f + df, r - 1)) bp |= 1ULL << ((r - 1) * 8 + f + df);
      }
    }
    ORC_KNIGHT_ATTACKS[sq] = kn;
    ORC_KING_ATTACKS[sq] = kg;
    ORC_WHITE_PAWN_ATTACKS[sq] = wp;
    ORC_BLACK_PAWN_ATTACKS[sq] = bp;
  }
}

/* Zobrist keys (generated.rs:7-27).  BLACK_TO_MOVE and the four castling keys are the
 * reference's constants (generated.rs:7-13).  The reference draws its piece and en passant
 * keys at random on every build (build.rs:26-39), so those values cannot be matched:
 * they are fixed here as consecutive splitmix64 outputs from the seed 0x70616269 ("pabi"),
 * 768 piece keys (index player * 384 + kind * 64 + square, generated.rs:22-27) then 8
 * file keys -- the recipe documented in include/pabi_cuda.h (pabi_zobrist_keys). */
static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static void build_keys(void) {
  uint64_t s = 0x70616269ULL; /* "pabi" */
  for (int i = 0; i < 768; i++) ORC_PIECE_KEYS[i] = splitmix64(&s);
  for (int i = 0; i < 8; i++) ORC_EP_KEYS[i] = splitmix64(&s);
  ORC_MISC_KEYS[0] = 0x9E06BAD39D761293ULL; /* BLACK_TO_MOVE, generated.rs:7 */
  ORC_MISC_KEYS[1] = 0xF05AC573DD61D323ULL; /* WHITE_CAN_CASTLE_SHORT, :9 */
  ORC_MISC_KEYS[2] = 0x41D8B55BA5FEB78BULL; /* WHITE_CAN_CASTLE_LONG, :10 */
  ORC_MISC_KEYS[3] = 0x680988787A43D289ULL; /* BLACK_CAN_CASTLE_SHORT, :12 */
  ORC_MISC_KEYS[4] = 0x2F941F8DFD3E3D1FULL; /* BLACK_CAN_CASTLE_LONG, :13 */
}

static int g_ready;

__attribute__((constructor)) void orc_init_tables(void) {
  if (g_ready) return;
  build_leapers();
  build_slider(BISHOP_DIRS, ORC_BISHOP_MASKS, ORC_BISHOP_OFFSETS, ORC_BISHOP_ATTACKS,
               ORC_BISHOP_ATTACKS_COUNT);
  build_slider(ROOK_DIRS, ORC_ROOK_MASKS, ORC_ROOK_OFFSETS, ORC_ROOK_ATTACKS,
               ORC_ROOK_ATTACKS_COUNT);
  build_rays();
  build_keys();
  g_ready = 1;
}

const uint64_t *oracle_table(const char *name, size_t *count) {
  static const struct {
    const char *name;
    const uint64_t *data;
    size_t n;
  } T[] = {
      {"knight_attacks", ORC_KNIGHT_ATTACKS, 64},
      {"king_attacks", ORC_KING_ATTACKS, 64},
      {"white_pawn_attacks", ORC_WHITE_PAWN_ATTACKS, 64},
      {"black_pawn_attacks", ORC_BLACK_PAWN_ATTACKS, 64},
      {"bishop_relevant_occupancies", ORC_BISHOP_MASKS, 64},
      {"rook_relevant_occupancies", ORC_ROOK_MASKS, 64},
      {"bishop_attack_offsets", ORC_BISHOP_OFFSETS, 64},
      {"rook_attack_offsets", ORC_ROOK_OFFSETS, 64},
      {"bishop_attacks", ORC_BISHOP_ATTACKS, ORC_BISHOP_ATTACKS_COUNT},
      {"rook_attacks", ORC_ROOK_ATTACKS, ORC_ROOK_ATTACKS_COUNT},
      {"rays", ORC_RAYS, 4096},
      {"bishop_rays", ORC_BISHOP_RAYS, 4096},
      {"rook_rays", ORC_ROOK_RAYS, 4096},
  };
  orc_init_tables();
  for (size_t i = 0; i < sizeof(T) / sizeof(T[0]); i++)
    if (strcmp(T[i].name, name) == 0) {
      if (count) *count = T[i].n;
      return T[i].data;
    }
  if (count) *count = 0;
  return NULL;
}

uint64_t oracle_rook_attacks(int sq, uint64_t occ) { return orc_rook_attacks(sq, occ); }
uint64_t oracle_bishop_attacks(int sq, uint64_t occ) { return orc_bishop_attacks(sq, occ); }
uint64_t oracle_ray(int from, int to) { return orc_ray(from, to); }

int oracle_has_bmi2(void) {
#ifdef __BMI2__
  return 1;
#else
  return 0;
#endif
}

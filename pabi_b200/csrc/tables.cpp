/*
 * tables.cpp -- builds the lookup tables from their definitions.
 *
 * What pabi ships as generated Rust consts (generated.rs:30-82) is rebuilt here:
 *   KNIGHT/KING attacks              generated/{knight,king}_attacks.rs
 *   *_RELEVANT_OCCUPANCIES           ray squares minus the last square of each ray
 *   *_ATTACK_OFFSETS                 prefix sums of 1 << popcount(mask)  (5 248 / 102 400 total)
 *   BISHOP_ATTACKS / ROOK_ATTACKS    ray-walk attack sets incl. the first blocker
 *   RAYS                             [from, to) on a common line, else empty
 * The per-square blocks of the slider tables keep pabi's offsets and sizes; the
 * position of an entry inside its block is given by a multiply-shift hash of the
 * relevant occupancy rather than PEXT (attacks.rs:67-87), because sm_100a has no
 * PEXT instruction and a bit loop would cost ~12 iterations per lookup.  The
 * multipliers are found here by a seeded search and verified exhaustively.
 */
#include "tables.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pabi {

static bool on_board(int f, int r) { return f >= 0 && f < 8 && r >= 0 && r < 8; }

static u64 walk(int sq, u64 occ, const int (*dirs)[2]) {
  u64 att = 0;
  for (int d = 0; d < 4; d++) {
    int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
    while (on_board(f, r)) {
      att |= bit(r * 8 + f);
      if (occ & bit(r * 8 + f)) break;
      f += dirs[d][0], r += dirs[d][1];
    }
  }
  return att;
}
static const int BISHOP_DIRS[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
static const int ROOK_DIRS[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
u64 walk_bishop_attacks(int sq, u64 occ) { return walk(sq, occ, BISHOP_DIRS); }
u64 walk_rook_attacks(int sq, u64 occ) { return walk(sq, occ, ROOK_DIRS); }

/* Relevant occupancy: every ray square except the last one of each ray. */
static u64 relevant_mask(int sq, const int (*dirs)[2]) {
  u64 m = 0;
  for (int d = 0; d < 4; d++) {
    int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
    while (on_board(f + dirs[d][0], r + dirs[d][1])) {
      m |= bit(r * 8 + f);
      f += dirs[d][0], r += dirs[d][1];
    }
  }
  return m;
}

static u64 leaper(int sq, const int (*steps)[2], int n) {
  u64 att = 0;
  for (int i = 0; i < n; i++) {
    int f = (sq & 7) + steps[i][0], r = (sq >> 3) + steps[i][1];
    if (on_board(f, r)) att |= bit(r * 8 + f);
  }
  return att;
}

struct Rng { /* xorshift64*, fixed seed: the tables are identical in every process */
  u64 s;
  u64 next() {
    s ^= s >> 12, s ^= s << 25, s ^= s >> 27;
    return s * 0x2545F4914F6CDD1DULL;
  }
  u64 sparse() { return next() & next() & next(); }
};

/* Fills table[offset .. offset + 2^bits) and returns the multiplier. */
static u64 build_slider(int sq, const int (*dirs)[2], u64 mask, u64* block, Rng& rng) {
  const int bits = popc(mask);
  const u32 size = 1u << bits;
  static u64 occs[4096], atts[4096];
  u64 sub = 0;
  for (u32 i = 0; i < size; i++) { /* carry-rippler over all subsets of mask */
    occs[i] = sub;
    atts[i] = walk(sq, sub, dirs);
    sub = (sub - mask) & mask;
  }
  static u32 epoch[4096];
  static u32 stamp = 0;
  for (int attempt = 0; attempt < 100000000; attempt++) {
    const u64 magic = rng.sparse();
    if (popc((mask * magic) >> 56) < 6) continue;
    ++stamp;
    bool ok = true;
    for (u32 i = 0; i < size && ok; i++) {
      const u32 idx = (u32)((occs[i] * magic) >> (64 - bits));
      if (epoch[idx] != stamp) {
        epoch[idx] = stamp;
        block[idx] = atts[i];
      } else if (block[idx] != atts[i]) {
        ok = false;
      }
    }
    if (ok) {
      /* slots no occupancy maps to are never read; keep them deterministic */
      for (u32 i = 0; i < size; i++)
        if (epoch[i] != stamp) block[i] = 0;
      return magic;
    }
  }
  std::fprintf(stderr, "pabi_b200: no slider multiplier found for square %d\n", sq);
  std::abort();
}

static HostTables* build() {
  HostTables* t = new HostTables;
  std::memset(t, 0, sizeof *t);
  SharedTables& s = t->shared;
  static const int KNIGHT_STEPS[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
  static const int KING_STEPS[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
  for (int sq = 0; sq < 64; sq++) {
    s.knight[sq] = leaper(sq, KNIGHT_STEPS, 8);
    s.king[sq] = leaper(sq, KING_STEPS, 8);
    u64 d = bit(sq), a = bit(sq);
    for (int k = 1; k < 8; k++) {
      int f = sq & 7, r = sq >> 3;
      if (on_board(f + k, r + k)) d |= bit((r + k) * 8 + f + k);
      if (on_board(f - k, r - k)) d |= bit((r - k) * 8 + f - k);
      if (on_board(f - k, r + k)) a |= bit((r + k) * 8 + f - k);
      if (on_board(f + k, r - k)) a |= bit((r - k) * 8 + f + k);
    }
    s.diag[sq] = d;
    s.anti[sq] = a;
    s.orth[sq] = (FILE_A << (sq & 7)) | (RANK_1 << (sq & 56));
    s.keep_from[sq] = s.keep_to[sq] = 15;
  }
  /* position.rs:435-468 (core.rs:604-612: 1 white short, 2 white long, 4 black short, 8 black long) */
  s.keep_from[4] = 15 & ~3, s.keep_from[7] = 15 & ~1, s.keep_from[0] = 15 & ~2;
  s.keep_from[60] = 15 & ~12, s.keep_from[63] = 15 & ~4, s.keep_from[56] = 15 & ~8;
  s.keep_to[7] = 15 & ~1, s.keep_to[0] = 15 & ~2, s.keep_to[63] = 15 & ~4, s.keep_to[56] = 15 & ~8;

  Rng rng{0x9E3779B97F4A7C15ULL};
  u32 boff = 0, roff = 0;
  for (int sq = 0; sq < 64; sq++) {
    const u64 bm = relevant_mask(sq, BISHOP_DIRS), rm = relevant_mask(sq, ROOK_DIRS);
    s.bishop[sq].mask = bm;
    s.bishop[sq].magic = build_slider(sq, BISHOP_DIRS, bm, s.bishop_attacks + boff, rng);
    s.bishop_os[sq] = SliderIndex{boff, (u32)(32 - popc(bm))};
    boff += 1u << popc(bm);
    s.rook[sq].mask = rm;
    s.rook[sq].magic = build_slider(sq, ROOK_DIRS, rm, t->rook_attacks + roff, rng);
    s.rook_os[sq] = SliderIndex{roff, (u32)(32 - popc(rm))};
    roff += 1u << popc(rm);
  }
  if (boff != BISHOP_TABLE_SIZE || roff != ROOK_TABLE_SIZE) {
    std::fprintf(stderr, "pabi_b200: slider table sizes %u/%u differ from generated.rs:30,44\n", boff, roff);
    std::abort();
  }
  for (int k = 0; k < 64; k++) { /* zones: the king's neighbours, plus the castling walk for a king on e1 / e8 */
    const u64 zone = s.king[k] | (k == 4 ? 0x6CULL : k == 60 ? 0x6CULL << 56 : 0);
    for (u64 z = zone; z; z &= z - 1) {
      const int t = lsb(z);
      s.knight_zone[k] |= s.knight[t];
      s.diag_zone[k] |= s.diag[t] | s.anti[t];
      s.orth_zone[k] |= (FILE_A << (t & 7)) | (RANK_1 << (t & 56));
    }
  }
  /* RAYS (attacks.rs:486-553): from inclusive, towards `to`, `to` exclusive */
  for (int from = 0; from < 64; from++)
    for (int to = 0; to < 64; to++) {
      if (from == to) continue;
      int df = (to & 7) - (from & 7), dr = (to >> 3) - (from >> 3);
      if (!(df == 0 || dr == 0 || df == dr || df == -dr)) continue;
      int sf = (df > 0) - (df < 0), sr = (dr > 0) - (dr < 0);
      u64 ray = 0;
      for (int f = from & 7, r = from >> 3; r * 8 + f != to; f += sf, r += sr) ray |= bit(r * 8 + f);
      t->rays[from * 64 + to] = ray;
    }
  { /* Zobrist keys: fixed (the reference re-draws the piece and file keys on every build, build.rs:26-39) */
    u64 state = ZOBRIST_SEED;
    auto splitmix64 = [&state] {
      u64 z = (state += 0x9E3779B97F4A7C15ULL);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      return z ^ (z >> 31);
    };
    for (int i = 0; i < ZOBRIST_BLACK_TO_MOVE; i++) t->zobrist[i] = splitmix64(); /* 768 piece keys, then 8 en passant files */
    t->zobrist[ZOBRIST_BLACK_TO_MOVE] = 0x9E06BAD39D761293ULL; /* generated.rs:7 */
    t->zobrist[ZOBRIST_CASTLE + 0] = 0xF05AC573DD61D323ULL;    /* WHITE_CAN_CASTLE_SHORT, generated.rs:9 */
    t->zobrist[ZOBRIST_CASTLE + 1] = 0x41D8B55BA5FEB78BULL;    /* WHITE_CAN_CASTLE_LONG, generated.rs:10 */
    t->zobrist[ZOBRIST_CASTLE + 2] = 0x680988787A43D289ULL;    /* BLACK_CAN_CASTLE_SHORT, generated.rs:12 */
    t->zobrist[ZOBRIST_CASTLE + 3] = 0x2F941F8DFD3E3D1FULL;    /* BLACK_CAN_CASTLE_LONG, generated.rs:13 */
  }
  return t;
}

u64 host_compute_hash(const Pos& p) {
  const HostTables& h = host_tables();
  const Tables tb{&h.shared, GlobalTables{h.rook_attacks, h.rays, h.zobrist}};
  return compute_hash(p, tb);
}

const HostTables& host_tables() {
  static HostTables* tables = nullptr;
  static std::once_flag once;
  std::call_once(once, [] { tables = build(); });
  return *tables;
}

} /* namespace pabi */

/*
 * kernels.cuh -- sm_100a kernels of the perft / move-generation path.
 *
 * K0  stage_tables        (device fn) 48 KB table image -> shared memory by one TMA bulk copy per block
 * K1  k_tile<TILE_EXPAND> one breadth-first ply: parents (SoA, HBM) -> children (SoA, HBM)
 * K2  k_tile<TILE_LEAF>   last two plies fused: parents -> children in shared memory ->
 *                         count-only move generation per child, nothing written to HBM
 *     k_count             generate_moves().len() per record (perft depth 1, sizing pass of K1/K3)
 * K3  k_tile<TILE_MOVES>  ordered u16 move lists at CSR offsets (same tile pipeline, coalesced stores)
 * K4  k_scan_*            exclusive prefix sum u32 -> u64 (reduce / spine / down-sweep)
 * K5  per-root accumulation, fused into k_leaf2 / k_count_accumulate
 *
 * All of them are HBM/issue-bound integer kernels; no tensor cores (nothing here
 * is a contraction).  Position arithmetic is chess.cuh.
 */
#pragma once
#include <cuda_runtime.h>

#include "chess.cuh"
#include "fen.cuh"
#include "incremental.cuh"

namespace pabi {

/* ---- shared-memory table staging (K0) -----------------------------------
 * One elected thread arms an mbarrier with the image size and issues a single TMA bulk copy
 * (cp.async.bulk global -> shared, UBLKCP in SASS); every thread then waits on the barrier's phase.
 * The copy engine moves the 48 KB while the other threads are already free to set up. */
__device__ __forceinline__ void stage_tables(SharedTables* dst, const SharedTables* __restrict__ src) {
  static_assert(sizeof(SharedTables) % 16 == 0 && sizeof(SharedTables) < (1u << 20), "bulk copies are 16-byte granular; tx count < 2^20");
  __shared__ __align__(8) unsigned long long table_barrier;
  const u32 bar = (u32)__cvta_generic_to_shared(&table_barrier);
  const u32 dst_s = (u32)__cvta_generic_to_shared(dst);
  constexpr u32 BYTES = (u32)sizeof(SharedTables);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); /* make the init visible to the async proxy */
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src),
                 "r"(BYTES), "r"(bar)
                 : "memory");
  }
  u32 done = 0;
#ifdef PABI_DEBUG_TRAPS /* debug builds only: a trap is a sticky context error, and a healthy copy can be slow under a profiler */
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
  while (!done) { /* phase 0 completes when the bytes have landed */
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
#ifdef PABI_DEBUG_TRAPS
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 5000000000ULL) __trap(); /* five seconds of wall time for a 48 KB copy */
#endif
  }
}

struct DeviceTables {
  const SharedTables* image; /* global copy of the shared-memory image */
  GlobalTables g;
};

struct PositionImage { /* == pabi_position_t (include/pabi_cuda.h) */
  u64 white[6], black[6], hash;
  u16 fullmove;
  u8 castling, stm, halfmove, ep, pad[2];
};
static_assert(sizeof(PositionImage) == 112, "pabi_position_t image");

/* View of a frontier: records in word pairs, pair k of record i at rec[(k * stride + i) * 2] (load_pos / store_pos). */
struct Frontier {
  u64* rec;
  u32* roots; /* root index of each record (perft_batch), may be null when single-root */
  size_t stride;
  u64* mult;  /* WEIGHTED accounting (perft with transposition merging): how many paths reach the record */
};

/* How leaf counts are attributed */
enum Accounting {
  SINGLE_ROOT, /* one total */
  PER_ROOT,    /* nodes[root of the record] */
  WEIGHTED     /* one total, every record counted `mult` times */
};

/* ---- block-wide helpers -------------------------------------------------- */
/* Barrier over one thread group of a block (bar.sync with an id and a thread count): groups of a
 * block synchronise independently, so one group's scan/generate phase overlaps the other's
 * counting phase while both share a single copy of the tables in shared memory. */
template <int GROUP_THREADS>
__device__ __forceinline__ void group_sync(int group) {
  if (GROUP_THREADS == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(GROUP_THREADS) : "memory");
}

/* Exclusive scan over the THREADS threads of one group (tid = index inside the group). */
template <int THREADS, int GROUP_THREADS = 0, class T = u32>
__device__ __forceinline__ T block_exclusive_scan(T v, T* warp_sums /* [THREADS/32 + 1] */, T& total, int tid_in = -1, int group = 0) {
  const int tid = tid_in >= 0 ? tid_in : (int)threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  T inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  group_sync<GROUP_THREADS>(group);
  if (warp == 0) {
    constexpr int NW = THREADS / 32;
    T w = lane < NW ? warp_sums[lane] : 0;
    T winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < NW) warp_sums[lane] = winc - w;
    if (lane == NW - 1) warp_sums[NW] = winc;
  }
  group_sync<GROUP_THREADS>(group);
  total = warp_sums[THREADS / 32];
  const T r = warp_sums[warp] + inc - v;
  group_sync<GROUP_THREADS>(group); /* warp_sums may be reused by the caller's next scan */
  return r;
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
  return v;
}

/* One atomicAdd per block on *dst. */
template <int THREADS>
__device__ __forceinline__ void block_add_u64(u64 v, u64* scratch /* [THREADS/32] */, unsigned long long* dst) {
  v = warp_sum_u64(v);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    u64 w = threadIdx.x < THREADS / 32 ? scratch[threadIdx.x] : 0;
    w = warp_sum_u64(w);
    if (threadIdx.x == 0 && w) atomicAdd(dst, (unsigned long long)w);
  }
}

/* ---- simple thread-per-record kernels ------------------------------------ */
constexpr int FLAT_THREADS = 256;

/* counts[i] = generate_moves().len() of record first + i */
__global__ void __launch_bounds__(FLAT_THREADS)
k_count(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t first, size_t n, u32* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS)
    counts[i] = count_moves(load_pos(rec, stride, first + i), tb);
}

/* perft with one ply left: nodes[root] += generate_moves().len() (position.rs:914-916) */
template <Accounting ACC>
__global__ void __launch_bounds__(FLAT_THREADS)
k_count_accumulate(DeviceTables dt, Frontier f, size_t n, const unsigned long long* __restrict__ n_in, unsigned long long* __restrict__ nodes) {
  constexpr bool MULTI_ROOT = ACC == PER_ROOT;
  if (n_in) n = (size_t)*n_in; /* a frontier sized on the device */
  if ((size_t)blockIdx.x * FLAT_THREADS >= n) return;
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  __shared__ u64 red[FLAT_THREADS / 32];
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  u64 acc = 0;
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    const u32 c = count_moves(load_pos(f.rec, f.stride, i), tb);
    if (MULTI_ROOT) {
      if (c) atomicAdd(nodes + f.roots[i], (unsigned long long)c);
    } else {
      acc += ACC == WEIGHTED ? c * f.mult[i] : (u64)c;
    }
  }
  if (!MULTI_ROOT) block_add_u64<FLAT_THREADS>(acc, red, nodes);
}

/* in_check (position.rs:735-746) */
__global__ void __launch_bounds__(FLAT_THREADS)
k_in_check(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t n, u8* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    Analysis a;
    analyze(load_pos(rec, stride, i), tb, a);
    out[i] = a.checkers != 0;
  }
}

/* AttackInfo (attacks.rs:89-200): five bitboards per record */
__global__ void __launch_bounds__(FLAT_THREADS)
k_attack_info(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t n, u64* __restrict__ out /* [n][5] */) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    const AttackInfoOut a = attack_info(load_pos(rec, stride, i), tb);
    u64* o = out + 5 * i;
    o[0] = a.attacks, o[1] = a.checkers, o[2] = a.pins, o[3] = a.xrays, o[4] = a.safe_king_squares;
  }
}

/* out[i] = rec[i] after moves[i] (position.rs:414-433); out_hash[i] = the hash field of in_aos[i] with the XORs of make_move */
__global__ void __launch_bounds__(FLAT_THREADS)
k_make_moves(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t n, const u16* __restrict__ moves,
             const PositionImage* __restrict__ in_aos, u64* __restrict__ out, size_t out_stride, u64* __restrict__ out_hash) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    Pos p = load_pos(rec, stride, i);
    out_hash[i] = in_aos[i].hash ^ hash_delta(p, tb, moves[i]);
    make_move(p, tb, moves[i]);
    store_pos(out, out_stride, i, p);
  }
}

/* Position::hash of the children of one breadth-first ply (pabi_cuda_expand_batch): child k of parent i, in
 * generate_moves order, gets the parent's hash field with the XORs of make_move (position.rs:414-732) */
__global__ void __launch_bounds__(FLAT_THREADS)
k_child_hashes(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t n, const u64* __restrict__ prefix,
               const PositionImage* __restrict__ in_aos, u64* __restrict__ child_hash) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    const Pos p = load_pos(rec, stride, i);
    const u64 h = in_aos[i].hash;
    u64 at = prefix[i] - prefix[0];
    generate_moves(p, tb, [&](int from, int to, int promo, int) { child_hash[at++] = h ^ hash_delta(p, tb, pack_move(from, to, promo)); });
  }
}

/* CSR offsets (u32) from the u64 prefix */
__global__ void k_prefix_to_offsets(const u64* __restrict__ prefix, size_t n, u32* __restrict__ offsets) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (size_t)gridDim.x * blockDim.x)
    offsets[i] = (u32)(prefix[i] - prefix[0]);
}

/* ---- AoS (pabi_position_t, 112 B) <-> SoA record conversion -------------- */

__device__ __forceinline__ PositionImage image_of(const Pos& p, u64 hash) {
  PositionImage q;
  const u64 w = p.white;
  q.white[0] = p.kings & w, q.black[0] = p.kings & ~w;
  q.white[1] = p.queens & w, q.black[1] = p.queens & ~w;
  q.white[2] = p.rooks & w, q.black[2] = p.rooks & ~w;
  q.white[3] = p.bishops & w, q.black[3] = p.bishops & ~w;
  q.white[4] = p.knights & w, q.black[4] = p.knights & ~w;
  q.white[5] = p.pawns & w, q.black[5] = p.pawns & ~w;
  q.hash = hash;
  q.fullmove = (u16)meta_fullmove(p.meta);
  q.castling = (u8)meta_castling(p.meta);
  q.stm = (u8)meta_stm(p.meta);
  q.halfmove = (u8)meta_halfmove(p.meta);
  q.ep = (u8)meta_ep(p.meta);
  q.pad[0] = q.pad[1] = 0;
  return q;
}


/* A record that cannot be a position (not exactly one king per side, pawns on a back rank: what validate() always
 * rejects, position.rs:934-960) would index the tables out of range; it is replaced by bare kings and counted in
 * *bad, and the call that uploaded it fails with an argument error. */
__global__ void k_pack(const PositionImage* __restrict__ in, size_t n, u64* __restrict__ rec, size_t stride, size_t first,
                       u32* __restrict__ roots, u32* __restrict__ bad) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const PositionImage& q = in[i];
    Pos p;
    p.white = q.white[0] | q.white[1] | q.white[2] | q.white[3] | q.white[4] | q.white[5];
    p.kings = q.white[0] | q.black[0];
    p.queens = q.white[1] | q.black[1];
    p.rooks = q.white[2] | q.black[2];
    p.bishops = q.white[3] | q.black[3];
    p.knights = q.white[4] | q.black[4];
    p.pawns = q.white[5] | q.black[5];
    p.meta = make_meta(q.castling & 15, q.ep > 63 ? NO_SQUARE : q.ep, q.stm & 1, q.halfmove, q.fullmove);
    if (popc(q.white[0]) != 1 || popc(q.black[0]) != 1 || (p.pawns & (RANK_1 | RANK_8))) {
      atomicAdd(bad, 1u);
      p.white = p.kings = bit(0), p.kings |= bit(63);
      p.pawns = p.knights = p.bishops = p.rooks = p.queens = 0;
      p.meta = make_meta(0, NO_SQUARE, 0, 0, 1);
    }
    store_pos(rec, stride, first + i, p);
    if (roots) roots[first + i] = (u32)i;
  }
}

/* hash: Position::hash of every record (maintained by the caller's kernel), or null -> 0 */
__global__ void k_unpack(const u64* __restrict__ rec, size_t stride, size_t n, const u64* __restrict__ hash,
                         PositionImage* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = image_of(load_pos(rec, stride, i), hash ? hash[i] : 0);
}

__device__ __forceinline__ u64 splitmix64(u64& s) {
  u64 z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* Seeded uniform-random playouts, one thread per game: the position source of the datagen-style
 * workload (BASELINE.json configs[4]; the loop a src/datagen driver would run around
 * generate_moves / make_move, position.rs:317-433).  At each ply the move is
 * moves[splitmix64(seed) % n] in generate_moves order; the position after every stride-th ply is
 * recorded until there is no legal move, the halfmove clock reaches 100
 * (Position::halfmove_clock_expired, position.rs:750-752), max_plies, or the game's slots are used. */
__global__ void __launch_bounds__(FLAT_THREADS)
k_random_playouts(DeviceTables dt, const u64* __restrict__ rec, size_t stride_rec, size_t n_games, const PositionImage* __restrict__ starts_aos,
                  const u64* __restrict__ seeds, u32 max_plies, u32 stride, u32 per_game_cap, PositionImage* __restrict__ out,
                  u32* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t g = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; g < n_games; g += (size_t)gridDim.x * FLAT_THREADS) {
    Pos p = load_pos(rec, stride_rec, g);
    u64 seed = seeds[g];
    u64 hash = starts_aos[g].hash; /* maintained as make_move does, position.rs:414-732 */
    u32 written = 0;
    PositionImage* slot = out + g * per_game_cap;
    for (u32 ply = 0;; ply++) {
      if (ply % stride == 0) {
        if (written >= per_game_cap) break;
        slot[written++] = image_of(p, hash);
      }
      if (ply >= max_plies || meta_halfmove(p.meta) >= 100) break;
      const u32 n = count_moves(p, tb);
      if (n == 0) break;
      const u32 pick = (u32)(splitmix64(seed) % n);
      u32 k = 0;
      u16 chosen = 0;
      generate_moves(p, tb, [&](int from, int to, int promo, int) {
        if (k++ == pick) chosen = pack_move(from, to, promo);
      });
      hash ^= hash_delta(p, tb, chosen);
      make_move(p, tb, chosen);
    }
    counts[g] = written;
  }
}

/* FEN / EPD lines -> positions, one thread per line (fen.cuh); line i is text[offsets[i] .. offsets[i+1]) */
__global__ void __launch_bounds__(FLAT_THREADS)
k_parse_fens(DeviceTables dt, const char* __restrict__ text, const u64* __restrict__ offsets, size_t n, PositionImage* __restrict__ out,
             signed char* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t i = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * FLAT_THREADS) {
    ParsedPosition p;
    size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    const char* line = text + offsets[i];
    if (len && line[len - 1] == '\n') --len; /* a line may carry its terminator */
    const int rc = parse_position(line, len, tb, p);
    status[i] = (signed char)rc;
    PositionImage q;
    for (int k = 0; k < 6; k++) q.white[k] = rc == 0 ? p.white[k] : 0, q.black[k] = rc == 0 ? p.black[k] : 0;
    q.hash = 0;
    if (rc == 0) q.hash = compute_hash(record_of(p), tb); /* from_fen caches compute_hash(), position.rs:266 */
    q.fullmove = rc == 0 ? p.fullmove : 0, q.castling = rc == 0 ? p.castling : 0, q.stm = rc == 0 ? p.stm : 0;
    q.halfmove = rc == 0 ? p.halfmove : 0, q.ep = rc == 0 ? p.ep : 0, q.pad[0] = q.pad[1] = 0;
    out[i] = q;
  }
}

/* ---- Game environment (game.rs:21-111 without the Syzygy branch), one thread per game ------- */
struct GamesState {
  u64* rec;          /* SoA records, stride n */
  u64* keys;         /* keys[g * (history_cap + 1) + k]: Position::hash after k actions (RepetitionTable, zobrist.rs:10-36) */
  u32* plies;        /* actions applied so far */
  u8* perspective;   /* side to move at the root (game.rs:36) */
  u8* threefold;     /* RepetitionTable::record() of the last apply (game.rs:56) */
  u32* overflow;     /* number of applies refused because the history was full */
  size_t n;
  u32 history_cap;
};

/* RepetitionTable::record (zobrist.rs:28-32): true exactly when the key has now been seen 3 times */
__device__ __forceinline__ bool record_key(const GamesState& gs, size_t g, u32 ply_after, u64 key) {
  u64* keys = gs.keys + g * ((size_t)gs.history_cap + 1);
  int count = 1;
  for (u32 k = 0; k < ply_after; k++) count += keys[k] == key;
  keys[ply_after] = key;
  return count == 3;
}

/* Game::result (game.rs:61-110): 0 none, 1 win, 2 draw, 3 loss, from the root player's perspective */
template <class TB>
__device__ __forceinline__ int game_result(const Pos& p, const TB& tb, bool threefold, int perspective, u32 n_moves) {
  if (threefold) return 2;
  if (meta_halfmove(p.meta) >= 100) return 2; /* Position::halfmove_clock_expired, position.rs:750-752 */
  if (n_moves == 0) {
    Analysis a;
    analyze(p, tb, a);
    if (a.checkers == 0) return 2;                       /* stalemate */
    return perspective == meta_stm(p.meta) ? 3 : 1;      /* the player to move is mated */
  }
  return 0;
}

__global__ void k_games_init(GamesState gs, const PositionImage* __restrict__ roots_aos) { /* Game::new, game.rs:30-47 */
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < gs.n; g += (size_t)gridDim.x * blockDim.x) {
    const Pos p = load_pos(gs.rec, gs.n, g);
    gs.keys[g * ((size_t)gs.history_cap + 1)] = roots_aos[g].hash; /* repetitions.record(root.hash()), game.rs:32 */
    gs.plies[g] = 0;
    gs.perspective[g] = (u8)meta_stm(p.meta);
    gs.threefold[g] = 0;
  }
}

/* the current Position of every game; its hash is the key recorded last (game.rs:56) */
__global__ void k_games_unpack(GamesState gs, PositionImage* __restrict__ out) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < gs.n; g += (size_t)gridDim.x * blockDim.x)
    out[g] = image_of(load_pos(gs.rec, gs.n, g), gs.keys[g * ((size_t)gs.history_cap + 1) + gs.plies[g]]);
}

/* Game::apply (game.rs:54-59); action 0xFFFF = leave the game untouched */
__global__ void __launch_bounds__(FLAT_THREADS) k_games_apply(DeviceTables dt, GamesState gs, const u16* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t g = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; g < gs.n; g += (size_t)gridDim.x * FLAT_THREADS) {
    const u16 action = actions[g];
    if (action == 0xFFFF) continue;
    const u32 ply = gs.plies[g];
    if (ply >= gs.history_cap) {
      atomicAdd(gs.overflow, 1u);
      continue;
    }
    Pos p = load_pos(gs.rec, gs.n, g);
    const u64 hash = gs.keys[g * ((size_t)gs.history_cap + 1) + ply] ^ hash_delta(p, tb, action);
    make_move(p, tb, action);
    store_pos(gs.rec, gs.n, g, p);
    gs.threefold[g] = record_key(gs, g, ply + 1, hash);
    gs.plies[g] = ply + 1;
  }
}

__global__ void __launch_bounds__(FLAT_THREADS) k_games_result(DeviceTables dt, GamesState gs, u8* __restrict__ results) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t g = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; g < gs.n; g += (size_t)gridDim.x * FLAT_THREADS) {
    const Pos p = load_pos(gs.rec, gs.n, g);
    results[g] = (u8)game_result(p, tb, gs.threefold[g] != 0, gs.perspective[g], count_moves(p, tb));
  }
}

/* Uniformly random self-play: actions()[splitmix64(seed) % n] until result() is Some or max_plies actions */
__global__ void __launch_bounds__(FLAT_THREADS)
k_games_play_random(DeviceTables dt, GamesState gs, const u64* __restrict__ seeds, u32 max_plies, u8* __restrict__ results,
                    u32* __restrict__ applied) {
  extern __shared__ __align__(16) unsigned char smem[];
  SharedTables* st = reinterpret_cast<SharedTables*>(smem);
  stage_tables(st, dt.image);
  const Tables tb{st, dt.g};
  for (size_t g = (size_t)blockIdx.x * FLAT_THREADS + threadIdx.x; g < gs.n; g += (size_t)gridDim.x * FLAT_THREADS) {
    Pos p = load_pos(gs.rec, gs.n, g);
    u64 seed = seeds[g];
    u32 ply = gs.plies[g], done = 0;
    u64 hash = gs.keys[g * ((size_t)gs.history_cap + 1) + ply];
    bool threefold = gs.threefold[g] != 0;
    const int perspective = gs.perspective[g];
    int r;
    for (;;) {
      const u32 n = count_moves(p, tb);
      r = game_result(p, tb, threefold, perspective, n);
      if (r != 0 || done >= max_plies) break;
      if (ply >= gs.history_cap) {
        atomicAdd(gs.overflow, 1u);
        break;
      }
      const u32 pick = (u32)(splitmix64(seed) % n);
      u32 k = 0;
      u16 chosen = 0;
      generate_moves(p, tb, [&](int from, int to, int promo, int) {
        if (k++ == pick) chosen = pack_move(from, to, promo);
      });
      hash ^= hash_delta(p, tb, chosen);
      make_move(p, tb, chosen);
      threefold = record_key(gs, g, ++ply, hash);
      ++done;
    }
    store_pos(gs.rec, gs.n, g, p);
    gs.plies[g] = ply;
    gs.threefold[g] = threefold;
    results[g] = (u8)r;
    applied[g] = done;
  }
}

/* Root-subtree sharding: the records of a frontier whose content hash is congruent to shard_index modulo shard_count.
 * The partition depends on the records only, not on their order in the frontier, so ranks that each expanded the tree
 * themselves (children land in no fixed order, TILE_EXPAND_ATOMIC) still pick disjoint, exhaustive shards.  Compaction
 * with one warp-aggregated atomicAdd per warp; *n_out is the shard's size. */
PB_HD u64 shard_hash(const Pos& p) {
  u64 h = 0x9E3779B97F4A7C15ULL;
  const u64 words[8] = {p.white, p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings, p.meta};
  for (int i = 0; i < 8; i++) {
    h = (h ^ words[i]) * 0xBF58476D1CE4E5B9ULL;
    h ^= h >> 29;
  }
  h *= 0x94D049BB133111EBULL;
  return h ^ (h >> 32);
}

__global__ void k_filter_shard(Frontier in, size_t n, const unsigned long long* __restrict__ n_in, u32 shard_index, u32 shard_count,
                               Frontier out, unsigned long long* __restrict__ n_out) {
  if (n_in) n = (size_t)*n_in;
  const size_t rounds = (n + (size_t)gridDim.x * blockDim.x - 1) / ((size_t)gridDim.x * blockDim.x);
  for (size_t r = 0; r < rounds; r++) { /* whole warps stay in the loop: the ballot below needs every lane */
    const size_t i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    Pos p;
    bool keep = false;
    if (i < n) {
      p = load_pos(in.rec, in.stride, i);
      keep = shard_hash(p) % shard_count == shard_index;
    }
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
    if (!mask) continue;
    const int lane = threadIdx.x & 31, leader = __ffs((int)mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(n_out, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (keep) {
      const size_t o = (size_t)base + __popc(mask & ((1u << lane) - 1));
      store_pos(out.rec, out.stride, o, p);
      if (out.roots) out.roots[o] = in.roots ? in.roots[i] : 0;
    }
  }
}

/* ---- K4: exclusive scan u32 -> u64, prefix[0..n] (prefix[n] = total) ------ */
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const u32* __restrict__ in, size_t n, u64* __restrict__ tile_sums) {
  __shared__ u64 red[SCAN_THREADS / 32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE;
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = warp_sum_u64(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    u64 w = threadIdx.x < SCAN_THREADS / 32 ? red[threadIdx.x] : 0;
    w = warp_sum_u64(w);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = w;
  }
}

/* single block: exclusive scan of the tile sums in place */
__global__ void __launch_bounds__(1024) k_scan_spine(u64* __restrict__ tile_sums, size_t n_tiles) {
  __shared__ u64 warp_tot[32];
  __shared__ u64 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t base = 0; base < n_tiles; base += 1024) {
    const size_t i = base + threadIdx.x;
    const u64 v = i < n_tiles ? tile_sums[i] : 0;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u64 o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u64 w = warp_tot[lane], winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        u64 o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
        if (lane >= d) winc += o;
      }
      warp_tot[lane] = winc - w;
    }
    __syncthreads();
    const u64 c = carry;
    if (i < n_tiles) tile_sums[i] = c + warp_tot[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + warp_tot[warp] + inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const u32* __restrict__ in, size_t n, const u64* __restrict__ tile_sums, u64* __restrict__ prefix) {
  __shared__ u32 warp_sums[SCAN_THREADS / 32 + 1];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = base + k < n ? in[base + k] : 0;
    s += v[k];
  }
  u32 total;
  u32 off = block_exclusive_scan<SCAN_THREADS>(s, warp_sums, total);
  u64 run = tile_sums[blockIdx.x] + off;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) prefix[base + k] = run;
    run += v[k];
    if (base + k + 1 == n) prefix[n] = run;
  }
}

/* ---- the first plies of a walk, one launch ------------------------------------------------
 * While a frontier has at most ROOT_THREADS records, one block expands it ply after ply (count,
 * block scan, make the moves, store the children in the reference's order) without returning to
 * the host: the launch-bound top of a perft tree (root -> 20 -> 400 -> 8 902 for the start
 * position) costs one launch instead of six per ply.  result = {plies done, size of the last
 * frontier written, records expanded}. */
constexpr int ROOT_THREADS = 1024;
constexpr int ROOT_MAX_PLIES = 8;
constexpr int ROOT_POOL = 16384; /* moves of one ply that the block makes cooperatively (a fuller ply: one thread per parent) */
struct RootLevels {
  Frontier level[ROOT_MAX_PLIES + 1];
};
struct RootShared {
  SharedTables tables;
  u64 parents[POS_WORDS][ROOT_THREADS];
  u32 pool[ROOT_POOL];
};

template <bool MULTI_ROOT>
__global__ void __launch_bounds__(ROOT_THREADS, 1)
k_root_levels(DeviceTables dt, RootLevels lv, u32 n, int max_plies, u64* __restrict__ result) {
  extern __shared__ __align__(16) unsigned char smem[];
  RootShared& rs = *reinterpret_cast<RootShared*>(smem);
  SharedTables* sh = &rs.tables;
  stage_tables(sh, dt.image);
  const Tables tb{sh, dt.g};
  __shared__ u32 warp_sums[ROOT_THREADS / 32 + 1];
  const int tid = threadIdx.x;
  int plies = 0;
  u64 expanded = 0;
  while (plies < max_plies && n && n <= (u32)ROOT_THREADS) {
    const Frontier cur = lv.level[plies], next = lv.level[plies + 1];
    Pos p;
    u32 cnt = 0, total;
    if ((u32)tid < n) {
      p = load_pos(cur.rec, cur.stride, tid);
      cnt = count_moves(p, tb);
      rs.parents[0][tid] = p.white, rs.parents[1][tid] = p.pawns, rs.parents[2][tid] = p.knights, rs.parents[3][tid] = p.bishops;
      rs.parents[4][tid] = p.rooks, rs.parents[5][tid] = p.queens, rs.parents[6][tid] = p.kings, rs.parents[7][tid] = p.meta;
    }
    u32 off = block_exclusive_scan<ROOT_THREADS>(cnt, warp_sums, total);
    if (total <= (u32)ROOT_POOL) {
      /* as in K1: the parents write their move lists into the pool, then the whole block makes the moves, one child
       * per thread and step, with coalesced stores (400 parents -> 8 902 children keep all 1 024 threads busy) */
      if (cnt) {
        u32 j = off;
        generate_moves(p, tb, [&](int from, int to, int promo, int kind) {
          if (PB_IN_BOUNDS(j, ROOT_POOL)) rs.pool[j] = ((u32)kind << 28) | ((u32)tid << 16) | pack_move(from, to, promo);
          ++j;
        });
      }
      __syncthreads();
      for (u32 c = tid; c < total; c += ROOT_THREADS) {
        const u32 e = rs.pool[c];
        const int parent = (e >> 16) & 0xFFF;
        Pos q;
        q.white = rs.parents[0][parent], q.pawns = rs.parents[1][parent], q.knights = rs.parents[2][parent];
        q.bishops = rs.parents[3][parent], q.rooks = rs.parents[4][parent], q.queens = rs.parents[5][parent];
        q.kings = rs.parents[6][parent], q.meta = rs.parents[7][parent];
        make_move_kind<-1, false, true>(q, tb, (u16)(e & 0xFFFF), (int)(e >> 28));
        store_pos(next.rec, next.stride, c, q);
        if (MULTI_ROOT) next.roots[c] = cur.roots[parent];
      }
    } else if (cnt) {
      const u32 root = MULTI_ROOT ? cur.roots[tid] : 0;
      generate_moves(p, tb, [&](int from, int to, int promo, int) {
        Pos q = p;
        make_move(q, tb, pack_move(from, to, promo));
        store_pos(next.rec, next.stride, off, q);
        if (MULTI_ROOT) next.roots[off] = root;
        ++off;
      });
    }
    expanded += n;
    n = total;
    ++plies;
    __syncthreads(); /* the children are the next ply's parents (same block: visible after the barrier) */
  }
  if (tid == 0) result[0] = (u64)plies, result[1] = n, result[2] = expanded;
}

/* Largest end in [start, n] with prefix[end] - prefix[start] <= capacity; result {end, children}. */
__global__ void k_chunk_end(const u64* __restrict__ prefix, size_t start, size_t n, u64 capacity, u64* __restrict__ result) {
  if (threadIdx.x || blockIdx.x) return;
  const u64 base = prefix[start];
  size_t lo = start, hi = n; /* invariant: prefix[lo] - base <= capacity */
  while (lo < hi) {
    const size_t mid = lo + (hi - lo + 1) / 2;
    if (prefix[mid] - base <= capacity) lo = mid; else hi = mid - 1;
  }
  result[0] = lo;
  result[1] = prefix[lo] - base;
}

/* ---- K3: ordered move lists, warp-synchronous ------------------------------------------------
 * Every warp owns 32 consecutive positions and a private pool in shared memory: each lane
 * generates its position's moves (reference order) into the pool at the offset the global scan
 * assigned, then the warp copies the pool to the CSR array with coalesced 64-byte stores.  No
 * block-wide barrier: move generation times differ a lot between unrelated positions, and with the
 * block-level tile pipeline 43 % of the stall samples were threads waiting for the slowest
 * generator of their group (ncu, round 1). */
constexpr int MOVES_THREADS = 1024;
constexpr int MOVES_POOL = 2048; /* u16 entries per warp; an average warp needs ~900, the worst case 32 * 218 */

struct MovesShared {
  SharedTables tables;
  u16 pool[MOVES_THREADS / 32][MOVES_POOL];
};

__global__ void __launch_bounds__(MOVES_THREADS, 1)
k_moves_warp(DeviceTables dt, const u64* __restrict__ rec, size_t stride, size_t n, const u64* __restrict__ prefix,
             u16* __restrict__ moves, u32* __restrict__ offsets) {
  extern __shared__ __align__(16) unsigned char smem[];
  MovesShared& sh = *reinterpret_cast<MovesShared*>(smem);
  stage_tables(&sh.tables, dt.image);
  const Tables tb{&sh.tables, dt.g};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u16* pool = sh.pool[warp];
  const size_t n_warps = (n + 31) / 32;
  for (size_t wi = (size_t)blockIdx.x * (MOVES_THREADS / 32) + warp; wi < n_warps; wi += (size_t)gridDim.x * (MOVES_THREADS / 32)) {
    const size_t first = wi * 32, i = first + lane;
    const bool valid = i < n;
    const u64 warp_prefix = prefix[first];
    const u32 total = (u32)(prefix[min(first + 32, n)] - warp_prefix);
    u32 off = 0, cnt = 0;
    Pos p;
    if (valid) {
      const u64 mine = prefix[i];
      off = (u32)(mine - warp_prefix);
      cnt = (u32)(prefix[i + 1] - mine);
      offsets[i] = (u32)(mine - prefix[0]);
      if (i + 1 == n) offsets[n] = (u32)(prefix[n] - prefix[0]);
      p = load_pos(rec, stride, i);
    }
    const u64 out_base = warp_prefix - prefix[0];
    for (u32 w = 0; w < total; w += MOVES_POOL) {
      if (cnt && off < w + MOVES_POOL && off + cnt > w) {
        u32 j = off;
        generate_moves(p, tb, [&](int from, int to, int promo, int) {
          const u32 slot = j++ - w; /* wraps to a huge value below the window */
          if (slot < (u32)MOVES_POOL) pool[slot] = pack_move(from, to, promo);
        });
      }
      __syncwarp();
      const u32 end = min(total - w, (u32)MOVES_POOL);
      for (u32 c = lane; c < end; c += 32) moves[out_base + w + c] = pool[c];
      __syncwarp();
    }
  }
}

/* ---- K3 fused: ordered move lists in ONE kernel ------------------------------------------------
 * generate_moves once per position, sizes scanned inside the kernel: no separate k_count pass, no k_scan_*
 * launches, no host round trip for the total, and the positions are read once -- straight from the caller's
 * 112-byte images (no k_pack) or from SoA records.  Everything is warp-synchronous (no block barrier: unrelated
 * positions take very different times to generate, see k_moves_warp):
 *   tile = the 32 positions of one warp pass, handed out by a ticket counter
 *   1  every lane generates its position's moves (reference order) into its column of the warp's
 *      staging area, at most LIST_LANE_CAP of them; the generator's return value is the count
 *   2  warp scan of the counts: where each list starts inside the tile
 *   3  decoupled look-back over the tiles before it, 32 status words at a time (aggregate / inclusive
 *      prefix), gives the tile's place in the whole batch; tiles of earlier launches of the same batch are
 *      complete, so a batch may be cut into chunks whose copies overlap the kernels
 *   4  offsets are written; the warp copies its staged lists to the CSR array with coalesced stores (the
 *      source lane of an output slot is found by a shuffle binary search over the lanes' offsets).  A warp
 *      holding a position with more than LIST_LANE_CAP moves regenerates into a sliding window instead.
 * Moves beyond moves_cap are not written; offsets[n] always holds the total, so the caller can retry. */
constexpr int LIST_THREADS = 1024;
#ifndef PABI_LIST_SLEEP_NS
#define PABI_LIST_SLEEP_NS 200
#endif
constexpr int LIST_LANE_CAP = 64; /* staged moves per position; the playout positions average 27.6, the record is 218 */

struct ListShared {
  SharedTables tables;
  u16 stage[LIST_THREADS / 32][32 * LIST_LANE_CAP]; /* [warp][k * 32 + lane]: lanes write neighbouring half-words */
};

struct ListScan {
  unsigned long long* status; /* one word per tile of the batch: bits 62-63 state (1 aggregate, 2 inclusive prefix), bits 0-61 moves */
  unsigned long long* ticket; /* zeroed before every launch */
  size_t tile_begin, tile_end; /* this launch covers these tiles of the batch */
};

struct ListInput {
  const PositionImage* aos; /* the caller's records, or null: */
  const u64* rec;           /* SoA records */
  size_t stride;
  u32* bad;                 /* aos: records that are not positions (see k_pack) */
};

__device__ __forceinline__ Pos list_position(const ListInput& in, size_t i) {
  if (!in.aos) return load_pos(in.rec, in.stride, i);
  const PositionImage& q = in.aos[i];
  Pos p;
  p.white = q.white[0] | q.white[1] | q.white[2] | q.white[3] | q.white[4] | q.white[5];
  p.kings = q.white[0] | q.black[0];
  p.queens = q.white[1] | q.black[1];
  p.rooks = q.white[2] | q.black[2];
  p.bishops = q.white[3] | q.black[3];
  p.knights = q.white[4] | q.black[4];
  p.pawns = q.white[5] | q.black[5];
  p.meta = make_meta(q.castling & 15, q.ep > 63 ? NO_SQUARE : q.ep, q.stm & 1, q.halfmove, q.fullmove);
  if (popc(q.white[0]) != 1 || popc(q.black[0]) != 1 || (p.pawns & (RANK_1 | RANK_8))) {
    atomicAdd(in.bad, 1u);
    p.white = p.kings = bit(0), p.kings |= bit(63);
    p.pawns = p.knights = p.bishops = p.rooks = p.queens = 0;
    p.meta = make_meta(0, NO_SQUARE, 0, 0, 1);
  }
  return p;
}

__global__ void __launch_bounds__(LIST_THREADS, 1)
k_moves_fused(DeviceTables dt, ListInput in, size_t n, ListScan scan, u16* __restrict__ moves, size_t moves_cap, u32* __restrict__ offsets) {
  extern __shared__ __align__(16) unsigned char smem[];
  ListShared& sh = *reinterpret_cast<ListShared*>(smem);
  if ((size_t)blockIdx.x * (LIST_THREADS / 32) >= scan.tile_end - scan.tile_begin) return;
  stage_tables(&sh.tables, dt.image);
  const Tables tb{&sh.tables, dt.g};
  const int lane = threadIdx.x & 31;
  u16* stage = sh.stage[threadIdx.x >> 5];
  constexpr unsigned long long VALUE = (1ULL << 62) - 1;
  for (;;) {
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(scan.ticket, 1ULL);
    const size_t tile = scan.tile_begin + (size_t)__shfl_sync(0xFFFFFFFFu, ticket, 0);
    if (tile >= scan.tile_end) break;
    const size_t i = tile * 32 + lane;
    const bool valid = i < n;
    Pos p;
    u32 cnt = 0;
    if (valid) { /* 1: the moves, staged */
      p = list_position(in, i);
      u32 k = 0;
      cnt = generate_moves(p, tb, [&](int from, int to, int promo, int) {
        if (k < (u32)LIST_LANE_CAP) stage[k * 32 + lane] = pack_move(from, to, promo);
        ++k;
      });
    }
    /* 2: where every list starts inside the tile */
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    const u32 off = inc - cnt, total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    const bool fat = __any_sync(0xFFFFFFFFu, cnt > (u32)LIST_LANE_CAP);
    /* 3: the tile's place in the batch */
    unsigned long long base = 0;
    if (tile > 0) {
      if (lane == 0) {
        __threadfence();
        atomicExch(scan.status + tile, (1ULL << 62) | total);
      }
      for (long long look = (long long)tile - 1;; look -= 32) { /* 32 predecessors per step, the nearest in lane 0 */
        const long long t = look - lane;
        unsigned long long s = 2ULL << 62; /* before the first tile: an inclusive prefix of zero */
        if (t >= 0) /* a predecessor that is still generating: sleep instead of spinning, the issue slots belong to the generators */
          while (((s = atomicAdd(scan.status + t, 0ULL)) >> 62) == 0) __nanosleep(PABI_LIST_SLEEP_NS);
        const unsigned inclusive = __ballot_sync(0xFFFFFFFFu, (s >> 62) == 2);
        const int stop = inclusive ? __ffs((int)inclusive) - 1 : 31; /* the nearest predecessor that knows its whole prefix */
        unsigned long long v = lane <= stop ? s & VALUE : 0;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        base += v;
        if (inclusive) break;
      }
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(scan.status + tile, (2ULL << 62) | (base + total));
    }
    if (valid) {
      offsets[i] = (u32)(base + off);
      if (i + 1 == n) offsets[n] = (u32)(base + off + cnt);
    }
    /* 4: the lists */
    if (base + total <= moves_cap) {
      if (!fat) {
        for (u32 c0 = 0; c0 < total; c0 += 32) { /* whole warps take every trip: the shuffles need all lanes */
          const u32 c = c0 + lane;
          int owner = 0; /* the last lane whose list starts at or before slot c */
#pragma unroll
          for (int step = 16; step; step >>= 1) {
            const int probe = owner + step;
            const u32 start = __shfl_sync(0xFFFFFFFFu, off, probe & 31);
            if (probe < 32 && start <= c) owner = probe;
          }
          const u32 first = __shfl_sync(0xFFFFFFFFu, off, owner);
          if (c < total) moves[base + c] = stage[(c - first) * 32 + owner];
        }
      } else { /* a list longer than a staging column: sliding windows over the warp's lists, regenerated */
        constexpr u32 WINDOW = 32 * LIST_LANE_CAP;
        for (u32 w = 0; w < total; w += WINDOW) {
          __syncwarp();
          if (cnt && off < w + WINDOW && off + cnt > w) {
            u32 j = off;
            generate_moves(p, tb, [&](int from, int to, int promo, int) {
              const u32 slot = j++ - w; /* wraps to a huge value below the window */
              if (slot < WINDOW) stage[slot] = pack_move(from, to, promo);
            });
          }
          __syncwarp();
          const u32 end = min(total - w, WINDOW);
          for (u32 c = lane; c < end; c += 32) moves[base + w + c] = stage[c];
        }
      }
    }
    __syncwarp(); /* the staging area is reused by the next tile */
  }
}

/* ---- transposition merging (perft "with hashing", made exact) -------------------------------
 * Records of one ply that are the same position (placement, side to move, castling rights, en
 * passant square; the clocks do not influence perft) are merged into one record whose multiplicity
 * is the sum.  A 64-bit tag table groups candidates; k_dedup_verify then compares the full records,
 * so a tag collision between different positions costs a missed merge, never a wrong count. */
PB_HD u64 dedup_tag(const Pos& p) {
  u64 h = 0x9E3779B97F4A7C15ULL;
  const u64 words[8] = {p.white, p.pawns, p.knights, p.bishops, p.rooks, p.queens, p.kings, p.meta & 0xFFFULL};
  for (int i = 0; i < 8; i++) {
    h = (h ^ words[i]) * 0xBF58476D1CE4E5B9ULL;
    h ^= h >> 29;
  }
  h *= 0x94D049BB133111EBULL;
  return (h ^ (h >> 32)) | 1; /* 0 marks an empty slot */
}

/* open addressing on the tag; the thread that claims a slot becomes the group's owner */
__global__ void k_dedup_insert(const u64* __restrict__ rec, size_t stride, size_t n, unsigned long long* __restrict__ tags, size_t mask,
                               u32* __restrict__ owner, u32* __restrict__ slot_of, u32* __restrict__ winner) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 tag = dedup_tag(load_pos(rec, stride, i));
    size_t slot = (size_t)(tag >> 1) & mask;
    for (;;) {
      const unsigned long long old = atomicCAS(tags + slot, 0ULL, (unsigned long long)tag);
      if (old == 0ULL) {
        owner[slot] = (u32)i;
        winner[i] = 1;
        break;
      }
      if (old == tag) {
        winner[i] = 0;
        break;
      }
      slot = (slot + 1) & mask;
    }
    slot_of[i] = (u32)slot;
  }
}

/* every non-owner compares itself with its group's owner: equal -> its multiplicity moves to the owner */
__global__ void k_dedup_verify(const u64* __restrict__ rec, size_t stride, size_t n, const u32* __restrict__ owner,
                               const u32* __restrict__ slot_of, u32* __restrict__ winner, unsigned long long* __restrict__ mult) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (winner[i]) continue;
    const size_t j = owner[slot_of[i]];
    const Pos a = load_pos(rec, stride, i), b = load_pos(rec, stride, j);
    const bool same = a.white == b.white && a.pawns == b.pawns && a.knights == b.knights && a.bishops == b.bishops &&
                      a.rooks == b.rooks && a.queens == b.queens && a.kings == b.kings && ((a.meta ^ b.meta) & 0xFFFULL) == 0;
    if (same) atomicAdd(mult + j, mult[i]);
    else winner[i] = 1; /* different positions with one tag: keep both */
  }
}

/* survivors move to out[prefix[i]] with their accumulated multiplicity */
__global__ void k_dedup_compact(const u64* __restrict__ rec, size_t stride, size_t n, const u32* __restrict__ winner,
                                const u64* __restrict__ prefix, const u64* __restrict__ mult, u64* __restrict__ out_rec,
                                size_t out_stride, u64* __restrict__ out_mult) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!winner[i]) continue;
    const size_t o = (size_t)prefix[i];
    store_pos(out_rec, out_stride, o, load_pos(rec, stride, i));
    out_mult[o] = mult[i];
  }
}

/* ---- K1 / K2: tile kernels ------------------------------------------------
 * A thread group takes PARENTS consecutive parents at a time:
 *   A   threads 0..PARENTS-1 write their parent's moves into the group's shared-memory pool as
 *       (kind << 28 | parent << 16 | move).
 *         K1 / K3: at the offsets of the global scan (k_count + k_scan_*), reference order,
 *             sliding windows when a tile has more moves than the pool
 *         K2: one enumeration per parent (enumerate_split, incremental.cuh): the INERT moves
 *             are settled on the spot (replies = inert x total), the slots of the clean and the
 *             other moves are reserved with a shared-memory atomic per piece (DirectEmitter):
 *             clean entries from the front of the pool, the others from its end; a range of
 *             parents that does not fit is halved and redone
 *   B   every thread takes entries from the pool, parent record from shared memory:
 *         K1: make_move in registers, store the child record (coalesced SoA) at its
 *             global offset
 *         K2: clean entry -> count_child_fast; other -> make_move_kind in registers and
 *             the full count; accumulate -- the children are never written to HBM
 * The geometry (threads, groups, parents per tile, pool entries) is a template parameter
 * so that variants can be timed against each other on the device (PABI_TILE_VARIANT,
 * see engine.cu). */
/* THREADS threads per block in GROUPS independent groups; each group works on tiles of PARENTS
 * parents with its own pool of POOL entries. */
template <int THREADS_, int PARENTS_, int POOL_, int MIN_BLOCKS_, int GROUPS_ = 1>
struct TileConfig {
  static constexpr int THREADS = THREADS_, PARENTS = PARENTS_, POOL = POOL_, MIN_BLOCKS = MIN_BLOCKS_, GROUPS = GROUPS_;
  static constexpr int GROUP_THREADS = THREADS_ / GROUPS_;
  static_assert(THREADS_ % (32 * GROUPS_) == 0 && GROUPS_ <= 15, "groups are whole warps; barrier ids 1..15");
  static_assert(PARENTS_ <= GROUP_THREADS && PARENTS_ <= 4096 && POOL_ >= 256, "a tile's parents need a thread each; the pool must hold one move list");
};

template <class CFG>
struct TileGroupShared {
  ulonglong2 parents[POS_WORDS / 2][CFG::PARENTS]; /* word pairs: a parent is read with four 128-bit loads */
  u32 pool[CFG::POOL];
  u32 parent_nodes[CFG::PARENTS]; /* K2, multi-root: nodes below each parent */
  u64 warp_sums[CFG::GROUP_THREADS / 32 + 1];
  unsigned long long next_tile; /* TILE_LEAF: the group's next tile (ticket counter), taken while the current tile is in phase A */
  u32 start_len, fullest, shrunk; /* TILE_LEAF: parents per pool pass of the next tile, and what decides it (thread 0 only) */
  u32 cursor;                   /* TILE_LEAF: pool entries reserved so far, other << 16 | clean (DirectEmitter) */
  u32 overflow;                 /* TILE_LEAF: a reservation did not fit: the range of parents is halved and redone */
};

template <class CFG>
struct TileShared {
  SharedTables tables;
  TileGroupShared<CFG> group[CFG::GROUPS];
  u64 red[CFG::THREADS / 32];
};

template <class CFG>
__device__ __forceinline__ Pos tile_parent(const TileGroupShared<CFG>& sh, int k) {
  Pos q;
  const ulonglong2 a = sh.parents[0][k], b = sh.parents[1][k], c = sh.parents[2][k], d = sh.parents[3][k];
  q.white = a.x, q.pawns = a.y, q.knights = b.x, q.bishops = b.y, q.rooks = c.x, q.queens = c.y, q.kings = d.x, q.meta = d.y;
  return q;
}

template <class CFG>
__device__ __forceinline__ void put_tile_parent(TileGroupShared<CFG>& sh, int k, const Pos& p) {
  sh.parents[0][k] = make_ulonglong2(p.white, p.pawns), sh.parents[1][k] = make_ulonglong2(p.knights, p.bishops);
  sh.parents[2][k] = make_ulonglong2(p.rooks, p.queens), sh.parents[3][k] = make_ulonglong2(p.kings, p.meta);
}

enum TileMode {
  TILE_LEAF,   /* K2: count the children, write nothing */
  TILE_EXPAND, /* K1: write the children as SoA records */
  TILE_MOVES,  /* K3: write the moves themselves (u16, CSR) */
  TILE_EXPAND_VIRTUAL, /* as TILE_EXPAND_ATOMIC, but the children are not made: a child is written as an 8-byte entry
                          (index of its parent record, kind of the moving piece, move), and K2 makes it itself.  The ply
                          below the last materialised one costs 8 instead of 64 bytes per position to write and to read */
  TILE_EXPAND_ATOMIC /* K1 without a sizing pass: a tile counts its own parents, scans the counts inside the thread group and
                        reserves its children's range with ONE atomicAdd on the next level's record counter.  The children of a
                        ply land in no particular tile order (perft does not care), there is no k_count / k_scan_* launch, and the
                        size of the next frontier exists only on the device: the host does not synchronise between plies */
};

struct TileOut {
  Frontier children;                /* TILE_EXPAND */
  u16* moves;                       /* TILE_MOVES: moves[prefix[i] - prefix[first] ...] */
  u32* offsets;                     /* TILE_MOVES: offsets[i] = prefix[first + i] - prefix[first] (n + 1 entries) */
  unsigned long long* nodes;        /* TILE_LEAF: per-root (or single) leaf counts */
  unsigned long long* visited;      /* TILE_LEAF statistics: children generated */
  unsigned long long* ticket;       /* TILE_LEAF: tiles are handed out dynamically (zeroed before the launch) */
  /* frontier sizes that live on the device (the host does not wait for a ply to learn how large the next one is): */
  const unsigned long long* n_in;   /* when set, the number of input records is *n_in, not the launch argument */
  unsigned long long* n_out;        /* TILE_EXPAND_ATOMIC: the children's record counter (zeroed before the launch) */
  unsigned long long* overflow;     /* TILE_EXPAND_ATOMIC: set when a tile's children would not fit out_cap (the host then redoes
                                       the call on the exact, chunked path) */
  size_t out_cap;
  /* a virtual frontier: entry = parent record index << 32 | kind << 16 | move */
  unsigned long long* virt_out;     /* TILE_EXPAND_VIRTUAL: where the entries go */
  const unsigned long long* virt;   /* TILE_LEAF: when set, parent i of the launch is entry virt[i] applied to record (entry >> 32) of `in` */
};

/* K2's enumerate_split sink, DirectEmitter, lives in incremental.cuh (host-checkable). */

/* In TILE_EXPAND / TILE_MOVES the per-parent counts come from the global scan (`prefix`, built by
 * k_count + k_scan_*), which also fixes where the tile's output starts; in TILE_LEAF they are
 * computed here and scanned within the block. */
template <TileMode MODE, Accounting ACC, class CFG>
__global__ void __launch_bounds__(CFG::THREADS, CFG::MIN_BLOCKS)
k_tile(DeviceTables dt, Frontier in, size_t first, size_t n, const u64* __restrict__ prefix, TileOut out) {
  constexpr bool MULTI_ROOT = ACC == PER_ROOT;
  constexpr int PARENTS = CFG::PARENTS, POOL = CFG::POOL, GROUPS = CFG::GROUPS, GT = CFG::GROUP_THREADS;
  constexpr int SYNC = GROUPS == 1 ? 0 : GT; /* 0: plain __syncthreads */
  constexpr bool LEAF = MODE == TILE_LEAF;
  constexpr bool ATOMIC = MODE == TILE_EXPAND_ATOMIC || MODE == TILE_EXPAND_VIRTUAL;
  if (out.n_in) n = (size_t)*out.n_in; /* a frontier sized on the device: the grid was launched for an upper bound */
  if ((size_t)blockIdx.x * GROUPS * PARENTS >= n) return; /* no tile for this block (the same for all its threads) */
  extern __shared__ __align__(16) unsigned char smem[];
  TileShared<CFG>& block = *reinterpret_cast<TileShared<CFG>*>(smem);
  stage_tables(&block.tables, dt.image);
  const Tables tb{&block.tables, dt.g};
  const int group = GROUPS == 1 ? 0 : threadIdx.x / GT;
  const int tid = GROUPS == 1 ? threadIdx.x : threadIdx.x % GT;
  TileGroupShared<CFG>& sh = block.group[group];
  u64 acc = 0;
  u32 inert_children = 0; /* K2 statistics: children that were never materialised */
  const size_t n_tiles = (n + PARENTS - 1) / PARENTS;

  /* K2 takes tiles from a ticket counter (their cost varies, and a static split leaves a tail of idle
   * groups at the end of the launch); K1/K3 tiles are cheap and uniform enough for a static stride */
  if (LEAF) { /* the ticket of a tile is taken while the tile before it is in phase A, so that its parents can be prefetched */
    /* start_len: how many parents of a tile share one pass through the pool.  The whole tile at first; a pass whose entries do
     * not fit halves it, and the NEXT tiles start from the halved length instead of overflowing again (in positions with few
     * inert moves four tiles in ten overflowed, each time throwing away the enumeration of all their parents); it doubles again
     * after a tile whose fullest pass used at most 35 % of the pool.  Kept by thread 0 in shared memory: written before the
     * first barrier of a tile, read by the others after it. */
    if (tid == 0) sh.next_tile = atomicAdd(out.ticket, 1ULL), sh.start_len = PARENTS, sh.fullest = 0, sh.shrunk = 0;
    group_sync<SYNC>(group);
  }
  for (size_t tile = (size_t)blockIdx.x * GROUPS + group;; tile += (size_t)gridDim.x * GROUPS) {
    if (LEAF) {
      tile = (size_t)sh.next_tile; /* thread 0 overwrites it only after the next barrier (below) */
    }
    if (tile >= n_tiles) break;
    const size_t i = tile * PARENTS + tid;
    const bool valid = tid < PARENTS && i < n;
    if (LEAF) {
      /* Every parent enumerates its moves once: context, split, inert moves counted, the rest written straight into
       * the pool at slots reserved with a shared atomic (DirectEmitter).  The parents of a tile are taken in ranges
       * [lo, lo + len); a range whose entries do not fit the pool is halved and redone (a single parent always fits). */
      static_assert(POOL >= 1024 && POOL <= 32768, "one parent's moves fit; the packed cursor has 16 bits per class");
      const u32 tile_parents = (u32)min((size_t)PARENTS, n - tile * PARENTS);
      u32 lo = 0, len = 0; /* len = 0: the first pass of the tile */
      while (lo < tile_parents) {
        if (tid == 0) sh.cursor = 0, sh.overflow = 0;
        group_sync<SYNC>(group);
        if (len == 0) {
          if (tid == 0) sh.next_tile = atomicAdd(out.ticket, 1ULL); /* the next tile, once per tile */
          len = sh.start_len;
        }
        const bool mine = valid && (u32)tid >= lo && (u32)tid < lo + len;
        u32 inert_nodes = 0, inert_moves = 0;
        if (mine) {
          Pos p;
          if (out.virt) { /* the parent is made here: its record's side to move normalised to white, then the move */
            const unsigned long long e = out.virt[first + i];
            p = load_pos(in.rec, in.stride, (size_t)(e >> 32));
            u16 mv = (u16)e;
            if (meta_stm(p.meta) == 1) p = mirror(p), mv ^= 56 | (56 << 6);
            make_move_kind<0, false>(p, tb, mv, (int)(e >> 16) & 7);
          } else {
            p = load_pos(in.rec, in.stride, first + i);
            if (meta_stm(p.meta) == 0) p = mirror(p);
          }
          const TContext ctx = t_context(p, tb);
          p.meta = (with_king_squares(p) & ~(0xFFFFULL << META_BASE)) | ((u64)ctx.base << META_BASE);
          put_tile_parent(sh, tid, p);
          DirectEmitter emit{sh.pool, &sh.cursor, &sh.overflow, (u32)POOL, (u32)tid << 16, ctx.zone};
          enumerate_split(p, tb, ctx.reach, ctx.guard, emit);
          inert_moves = emit.inert;
          inert_nodes = emit.inert * ctx.total;
        }
        group_sync<SYNC>(group);
        if (sh.overflow) { /* the same for every thread of the group: read after the barrier, reset after the next one */
#ifdef PABI_STATS
          if (tid == 0) atomicAdd(out.visited + 3, 1ULL);
#endif
          len = len > 1 ? len >> 1 : 1;
          if (tid == 0) sh.shrunk = 1;
          group_sync<SYNC>(group);
          continue;
        }
        if (lo == 0 && tid < PARENTS && (tid & 15) == 0) { /* the next tile's parents towards L2, one request per 128-byte line */
          const size_t next_first = (size_t)sh.next_tile * PARENTS + tid;
          if (next_first < n && out.virt) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(out.virt + first + next_first)); /* 16 entries are one 128-byte line */
          } else if (next_first < n) {
#pragma unroll
            for (int w = 0; w < POS_WORDS / 2; w++) { /* 16 records of a pair array are two 128-byte lines */
              asm volatile("prefetch.global.L2 [%0];" ::"l"(in.rec + ((size_t)w * in.stride + first + next_first) * 2));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(in.rec + ((size_t)w * in.stride + first + next_first) * 2 + 16));
            }
          }
        }
        inert_children += inert_moves;
        if (MULTI_ROOT) { if (mine) sh.parent_nodes[tid] = inert_nodes; }
        else if (mine) acc += ACC == WEIGHTED ? (u64)inert_nodes * in.mult[first + i] : (u64)inert_nodes;
#ifdef PABI_STATS
        if (tid == 0) atomicAdd(out.visited + 2, 1ULL), atomicMax(out.visited + 4, (unsigned long long)((sh.cursor & 0xFFFF) + (sh.cursor >> 16)));
#endif
        const u32 cursor = sh.cursor;
        const u32 clean_here = cursor & 0xFFFF, others = cursor >> 16, end = clean_here + others;
        if (tid == 0 && end) atomicAdd(out.visited, (unsigned long long)end), sh.fullest = max(sh.fullest, end); /* statistics: positions at the last ply but one */
        if (MULTI_ROOT) group_sync<SYNC>(group); /* the per-parent counters are set */
        for (u32 c = tid; c < end; c += GT) {
          const u32 slot = c < clean_here ? c : POOL - others + (c - clean_here);
          const u32 e = PB_IN_BOUNDS(slot, POOL) ? sh.pool[slot] : 0;
          const int parent = PB_IN_BOUNDS((e >> 16) & 0xFFF, PARENTS) ? (e >> 16) & 0xFFF : 0;
          Pos q = tile_parent(sh, parent);
          u32 k;
          if (c < clean_here) { /* quiet move off every line that matters: replies = base + pawns + king */
            k = count_child_fast(q, tb, (int)(e & 63), (int)((e >> 6) & 63), (int)(e >> 28));
          } else {
            make_move_kind<1, true>(q, tb, (u16)(e & 0xFFFF), (int)(e >> 28));
            k = count_child_white(q, tb);
          }
          if (MULTI_ROOT) atomicAdd(&sh.parent_nodes[parent], k);
          else acc += ACC == WEIGHTED ? k * in.mult[first + tile * PARENTS + parent] : (u64)k;
        }
        if (MULTI_ROOT) {
          group_sync<SYNC>(group);
          if (mine && sh.parent_nodes[tid])
            atomicAdd(out.nodes + in.roots[out.virt ? (size_t)(out.virt[first + i] >> 32) : first + i], (unsigned long long)sh.parent_nodes[tid]);
        }
        group_sync<SYNC>(group); /* the pool and the cursor are free again */
        lo += len;
      }
      if (tid == 0) { /* the next tile's starting length (thread 0 passes its first barrier only after this) */
        if (sh.shrunk) sh.start_len = len;
        else if (sh.start_len < (u32)PARENTS && sh.fullest * 20 <= (u32)POOL * 7) sh.start_len <<= 1;
        sh.shrunk = 0, sh.fullest = 0;
      }
      continue;
    }
    /* K1 / K3 */
    u32 cnt = 0, off = 0, total;
    u64 out_base = 0; /* global index of this tile's first child / move */
    if (valid) {
      const Pos p = load_pos(in.rec, in.stride, first + i);
      put_tile_parent(sh, tid, p);
      if (ATOMIC) cnt = count_moves(p, tb);
    }
    if (ATOMIC) {
      /* sizes from the tile itself: scan of the parents' counts inside the thread group, then one reservation */
      off = block_exclusive_scan<GT, SYNC, u32>(cnt, reinterpret_cast<u32*>(sh.warp_sums), total, tid, group);
      if (tid == 0) {
        unsigned long long base = total ? atomicAdd(out.n_out, (unsigned long long)total) : 0ULL;
        if (base + total > out.out_cap) {
          atomicExch(out.overflow, 1ULL);
          base = ~0ULL;
        }
        sh.next_tile = base;
      }
      group_sync<SYNC>(group);
      out_base = sh.next_tile;
      group_sync<SYNC>(group); /* thread 0 writes next_tile again in the next tile */
      if (out_base == ~0ULL) continue; /* the same for the whole group */
    } else {
      const size_t tile_first = first + tile * PARENTS;
      const size_t tile_end = first + min((tile + 1) * (size_t)PARENTS, n);
      const u64 tile_prefix = prefix[tile_first];
      out_base = tile_prefix - prefix[first];
      total = (u32)(prefix[tile_end] - tile_prefix);
      if (valid) {
        const u64 mine = prefix[first + i];
        off = (u32)(mine - tile_prefix);
        cnt = (u32)(prefix[first + i + 1] - mine);
        if (MODE == TILE_MOVES) {
          out.offsets[i] = (u32)(mine - prefix[first]);
          if (i + 1 == n) out.offsets[n] = (u32)(prefix[first + n] - prefix[first]);
        }
      }
      group_sync<SYNC>(group); /* the parents are in shared memory */
    }

    {
      for (u32 w = 0; w < total; w += POOL) { /* sliding windows over the tile's children */
        if (cnt && off < w + POOL && off + cnt > w) {
          u32 j = off;
          generate_moves(tile_parent(sh, tid), tb, [&](int from, int to, int promo, int kind) {
            const u32 slot = j++ - w; /* wraps to a huge value below the window */
            if (slot < (u32)POOL) sh.pool[slot] = ((u32)kind << 28) | ((u32)tid << 16) | pack_move(from, to, promo);
          });
        }
        group_sync<SYNC>(group);
        const u32 end = min(total - w, (u32)POOL);
        for (u32 c = tid; c < end; c += GT) {
          const u32 e = sh.pool[c];
          if (MODE == TILE_MOVES) {
            out.moves[out_base + w + c] = (u16)e;
            continue;
          }
          const int parent = PB_IN_BOUNDS((e >> 16) & 0xFFF, PARENTS) ? (e >> 16) & 0xFFF : 0;
          if (MODE == TILE_EXPAND_VIRTUAL) { /* the child stays a (parent, move) pair */
            if (PB_IN_BOUNDS(out_base + w + c, out.out_cap))
              out.virt_out[out_base + w + c] = ((unsigned long long)(first + tile * PARENTS + parent) << 32) | ((unsigned long long)(e >> 28) << 16) | (e & 0xFFFF);
            continue;
          }
          Pos q = tile_parent(sh, parent);
          make_move_kind<-1, false, true>(q, tb, (u16)(e & 0xFFFF), (int)(e >> 28)); /* == make_move, without the piece lookup */
          const size_t o = (size_t)(out_base + w + c);
          store_pos(out.children.rec, out.children.stride, o, q);
          if (MULTI_ROOT) out.children.roots[o] = in.roots[first + tile * PARENTS + parent];
          if (ACC == WEIGHTED) out.children.mult[o] = in.mult[first + tile * PARENTS + parent];
        }
        group_sync<SYNC>(group);
      }
    }
  }
  if (LEAF) { /* statistics: the inert children were visited too */
    u32 v = inert_children;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out.visited, (unsigned long long)v);
  }
  if (LEAF && !MULTI_ROOT) block_add_u64<CFG::THREADS>(acc, block.red, out.nodes);
}

/* ---- K2, warp-synchronous variant ------------------------------------------------------------
 * The same three phases as k_tile<TILE_LEAF>, but a tile is 32 parents owned by ONE warp: count +
 * warp shuffle scan, generation into the warp's private pool, then the 32 lanes walk the pool.  No
 * block-wide barrier, no idle group while the slowest parent of 256 finishes generating; tiles come
 * from the ticket counter.  Pool entries are 3 bytes: the move, and kind << 5 | parent.
 * Measured (B200, startpos perft(8), before the fast paths of incremental.cuh existed): 107.8 ms, the same as the
 * block-level kernel with four 256-thread groups and tickets (107.8 ms) -- once tiles are scheduled dynamically the barrier waits are hidden by the
 * other groups and both versions sit on the ALU-pipe limit of count_moves_white.  The block-level kernel
 * stays the default because it also serves K1 and the per-root / weighted accountings, and it alone got the
 * quiet-move fast paths afterwards (45.8 ms); this one is kept as the plain reference formulation of K2,
 * selectable with PABI_TILE_VARIANT=8. */
constexpr int LEAFW_THREADS = 1024;
constexpr int LEAFW_POOL = 1152; /* entries per warp; 32 parents average ~830 children in the start-position tree */

struct LeafWarpShared {
  u64 parents[POS_WORDS][32];
  u16 moves[LEAFW_POOL];
  u8 tags[LEAFW_POOL];
};
struct LeafWarpBlock {
  SharedTables tables;
  LeafWarpShared warp[LEAFW_THREADS / 32];
  u64 red[LEAFW_THREADS / 32];
};

template <Accounting ACC>
__global__ void __launch_bounds__(LEAFW_THREADS, 1)
k_leaf_warp(DeviceTables dt, Frontier in, size_t n, TileOut out) {
  if (out.n_in) n = (size_t)*out.n_in; /* a frontier sized on the device */
  extern __shared__ __align__(16) unsigned char smem[];
  LeafWarpBlock& block = *reinterpret_cast<LeafWarpBlock*>(smem);
  stage_tables(&block.tables, dt.image);
  const Tables tb{&block.tables, dt.g};
  const int lane = threadIdx.x & 31;
  LeafWarpShared& sh = block.warp[threadIdx.x >> 5];
  const size_t n_tiles = (n + 31) / 32;
  u64 acc = 0;
  for (;;) {
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(out.ticket, 1ULL);
    const size_t tile = (size_t)__shfl_sync(0xFFFFFFFFu, ticket, 0);
    if (tile >= n_tiles) break;
    const size_t i = tile * 32 + lane;
    const bool valid = i < n;
    u32 cnt = 0;
    if (valid) {
      Pos p = load_pos(in.rec, in.stride, i);
      cnt = count_moves(p, tb);
      if (meta_stm(p.meta) == 0) p = mirror(p); /* children are then "white to move" */
      p.meta = with_king_squares(p);
      sh.parents[0][lane] = p.white, sh.parents[1][lane] = p.pawns, sh.parents[2][lane] = p.knights;
      sh.parents[3][lane] = p.bishops, sh.parents[4][lane] = p.rooks, sh.parents[5][lane] = p.queens;
      sh.parents[6][lane] = p.kings, sh.parents[7][lane] = p.meta;
    }
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    const u32 off = inc - cnt;
    if (lane == 0 && total) atomicAdd(out.visited, (unsigned long long)total);
    u64 parent_acc = 0; /* PER_ROOT: nodes below this lane's parent */
    __syncwarp();
    for (u32 w = 0; w < total; w += LEAFW_POOL) {
      if (cnt && off < w + LEAFW_POOL && off + cnt > w) {
        Pos p;
        p.white = sh.parents[0][lane], p.pawns = sh.parents[1][lane], p.knights = sh.parents[2][lane], p.bishops = sh.parents[3][lane];
        p.rooks = sh.parents[4][lane], p.queens = sh.parents[5][lane], p.kings = sh.parents[6][lane], p.meta = sh.parents[7][lane];
        u32 j = off;
        generate_moves<1>(p, tb, [&](int from, int to, int promo, int kind) {
          const u32 slot = j++ - w; /* wraps to a huge value below the window */
          if (slot < (u32)LEAFW_POOL) sh.moves[slot] = pack_move(from, to, promo), sh.tags[slot] = (u8)(kind << 5 | lane);
        });
      }
      __syncwarp();
      const u32 end = min(total - w, (u32)LEAFW_POOL);
      for (u32 c = lane; c < end; c += 32) {
        const u32 tag = sh.tags[c];
        const int parent = tag & 31;
        Pos q;
        q.white = sh.parents[0][parent], q.pawns = sh.parents[1][parent], q.knights = sh.parents[2][parent];
        q.bishops = sh.parents[3][parent], q.rooks = sh.parents[4][parent], q.queens = sh.parents[5][parent];
        q.kings = sh.parents[6][parent], q.meta = sh.parents[7][parent];
        make_move_kind<1, true>(q, tb, sh.moves[c], (int)(tag >> 5));
        const u32 k = count_child_white(q, tb);
        if (ACC == PER_ROOT) atomicAdd(out.nodes + in.roots[tile * 32 + parent], (unsigned long long)k);
        else acc += ACC == WEIGHTED ? k * in.mult[tile * 32 + parent] : (u64)k;
      }
      __syncwarp();
    }
    (void)parent_acc;
  }
  if (ACC != PER_ROOT) block_add_u64<LEAFW_THREADS>(acc, block.red, out.nodes);
}

} /* namespace pabi */

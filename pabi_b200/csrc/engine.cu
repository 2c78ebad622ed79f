/*
 * engine.cu -- host driver of the device path and the `pabi_cuda_*` C ABI
 * (include/pabi_cuda.h).
 *
 * perft (position.rs:909-924) is organised as a chunked breadth-first search:
 * the plies above the last two are materialised level by level in HBM as SoA
 * frontiers (count -> scan -> expand, children in the reference's move order);
 * the last two plies run in the fused kernel k_leaf2 which never writes its
 * children out.  A level whose children would not fit the next level's buffer
 * is cut into chunks, each finished depth-first before the next starts, so
 * memory use is bounded by depth * capacity whatever the tree size.
 *
 * There is no CPU implementation of move generation in this library: without a
 * CUDA device every entry point here fails with the CUDA error code.
 */
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <map>
#include <set>
#include <thread>
#include <vector>

#include "../../include/pabi_cuda.h"
#include "kernels.cuh"
#include "tables.h"

using namespace pabi;

namespace {

constexpr size_t DEFAULT_CAPACITY = size_t(32) << 20; /* records per level buffer */
constexpr int MAX_LEVELS = 24;
constexpr int FAST_SLOTS = 2048;      /* device-sized frontiers per call (one counter each); the last slot is the overflow flag */
constexpr size_t BRANCH_ESTIMATE = 32; /* children per position assumed when a frontier's size is only known on the device */
constexpr size_t BRANCH_MARGIN = 4;   /* a ply is expanded without its exact size only if 4 x the estimate still fits the next buffer */

struct Level {
  u64* rec = nullptr;
  u32* roots = nullptr;
  u32* counts = nullptr;
  u64* prefix = nullptr;
  u64* mult = nullptr; /* only for perft with transposition merging */
  size_t cap = 0;
};

struct Buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
};

/* growable device scratch of a context; a call may use each slot for one thing */
enum Scratch {
  S_STAGING = 0, /* host positions (AoS) on their way in */
  S_SMALL_A,     /* offsets / seeds / flags / moves-in */
  S_SMALL_B,     /* moves-out / counts / records */
  S_AOS_OUT,     /* positions (AoS) on their way out */
  S_COUNTS,      /* u32 per record: move counts of the device-resident API, slots of the merge table */
  S_PREFIX,      /* u64 per record: scan of the device-resident API */
  S_SLOTS
};

} /* namespace */

struct pabi_cuda_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  size_t capacity = DEFAULT_CAPACITY;
  DeviceTables dt{};
  void* d_tables = nullptr;
  u64* d_rook = nullptr;
  u64* d_rays = nullptr;
  u64* d_zobrist = nullptr;
  u32* d_bad_roots = nullptr;      /* k_pack: records that are not positions (replaced by bare kings, the call fails) */
  u32* h_bad_roots = nullptr;      /* pinned copy, valid after the call's final synchronisation */
  std::set<const void*> big_smem;  /* tile-kernel instantiations whose dynamic shared-memory limit has been raised */
  /* frontier sizes that exist only on the device ("levels sized on the device"): one 64-bit counter per
   * device-sized frontier of a call, FAST_SLOTS - 1 the overflow flag; read back once, when the call ends */
  unsigned long long* d_fast = nullptr;
  unsigned long long* h_fast = nullptr; /* pinned */
  int fast_used = 0;                    /* counters handed out in this call */
  std::vector<int> fast_interior, fast_leaf; /* counters whose frontier was expanded / was the input of K2 (statistics) */
  struct FastEdge { size_t parent_n; int parent_slot, child_slot; }; /* a speculative ply: parent (exact size or counter) -> children's counter */
  std::vector<FastEdge> fast_edges;
  double branching = 0; /* children per position of the last speculative ply whose two sizes are known (0: none yet) */
  bool exact_only = false;              /* this call walks on the exact path (a speculative walk overflowed, or PABI_EXACT_LEVELS=1) */
  bool force_exact = false;
  bool no_virtual = false;              /* PABI_NO_VIRTUAL_PLY=1: the ply above K2 is always materialised (comparison runs) */
  /* move lists (K3 fused): a second and a third stream so that the copies of a batch overlap its kernels */
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> list_events; /* per chunk: positions on the device, kernel started, kernel finished */
  unsigned long long* d_list_status = nullptr; /* one scan word per tile of a batch */
  size_t list_status_cap = 0;
  unsigned long long* h_list_totals = nullptr; /* pinned: moves up to the end of every chunk */
  size_t list_totals_cap = 0;
  bool list_two_pass = false; /* PABI_LIST_TWO_PASS=1: the round-1 path (k_count, k_scan_*, k_moves_warp), kept for comparison */
  uint32_t shard_index = 0, shard_count = 1; /* this call keeps the subtrees of one shard ... */
  int shard_level = -1;                      /* ... cut at this level (-1: no sharding) */
  Level levels[MAX_LEVELS];
  Level spare; /* shard gather target */
  Level merge_level; /* compaction target of merge_transpositions */
  unsigned long long* d_tags = nullptr;
  u32* d_owner = nullptr;
  size_t merge_table = 0;
  uint64_t merged_away = 0; /* per call: records removed by transposition merging */
  u64* d_tile_sums = nullptr;
  size_t tile_sums_cap = 0;
  u64* d_result = nullptr; /* {end, children} of k_chunk_end */
  unsigned long long* d_visited = nullptr; /* statistics: children generated inside k_tile<LEAF> */
  u64* h_result = nullptr; /* pinned */
  unsigned long long* d_nodes = nullptr;
  size_t nodes_cap = 0;
  Buffer scratch[S_SLOTS];
  int tile_variant = 0;
  std::string error;
  /* per-call statistics */
  uint64_t launches = 0, interior = 0, leaf_launches = 0, leaf_parents = 0;
  std::vector<cudaEvent_t> leaf_events; /* begin/end pairs around every k_tile<LEAF> launch */
};

namespace {

struct CudaFail {
  cudaError_t code;
};

#define CK(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      ctx->error = std::string(#expr) + ": " + cudaGetErrorString(e_);                       \
      throw CudaFail{e_};                                                                     \
    }                                                                                         \
  } while (0)

#define LAUNCHED()                      \
  do {                                  \
    ++ctx->launches;                    \
    CK(cudaGetLastError());             \
  } while (0)

int grid_for(const pabi_cuda_ctx* ctx, size_t n, int threads, int per_sm) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t)ctx->sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  return blocks ? (int)blocks : 1;
}

void* ensure(pabi_cuda_ctx* ctx, int slot, size_t bytes) {
  Buffer& b = ctx->scratch[slot];
  if (b.bytes < bytes) {
    if (b.ptr) CK(cudaFree(b.ptr));
    b.ptr = nullptr, b.bytes = 0;
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&b.ptr, want));
    b.bytes = want;
  }
  return b.ptr;
}

void alloc_level(pabi_cuda_ctx* ctx, Level& l, size_t cap) {
  if (l.cap >= cap) return;
  if (l.rec) cudaFree(l.rec), cudaFree(l.roots), cudaFree(l.counts), cudaFree(l.prefix), cudaFree(l.mult);
  l = Level();
  CK(cudaMalloc(&l.rec, cap * POS_WORDS * sizeof(u64)));
  CK(cudaMalloc(&l.roots, cap * sizeof(u32)));
  CK(cudaMalloc(&l.counts, cap * sizeof(u32)));
  CK(cudaMalloc(&l.prefix, (cap + 1) * sizeof(u64)));
  l.cap = cap;
}

Frontier frontier_of(const Level& l) { return Frontier{l.rec, l.roots, l.cap, l.mult}; }

void ensure_mult(pabi_cuda_ctx* ctx, Level& l) {
  if (!l.mult) CK(cudaMalloc(&l.mult, l.cap * sizeof(u64)));
}

void ensure_nodes(pabi_cuda_ctx* ctx, size_t n) {
  if (ctx->nodes_cap >= n) return;
  if (ctx->d_nodes) CK(cudaFree(ctx->d_nodes));
  ctx->d_nodes = nullptr;
  CK(cudaMalloc(&ctx->d_nodes, n * sizeof(unsigned long long)));
  ctx->nodes_cap = n;
}

constexpr size_t FLAT_SMEM = sizeof(SharedTables);

/* Tile-kernel geometries.  Variant 0 is the default; the others exist so that the
 * choice can be re-measured on the device (environment variable PABI_TILE_VARIANT). */
/* 1024 threads (64 registers each, 32 warps per SM) in two independent 512-thread groups that share one table
 * image: while one group generates, the other counts.  The pool takes what is left of the 227 KB of shared memory
 * after the table image and the parents.  Measured on startpos perft(8) with one enumeration per parent (DirectEmitter):
 * 2 x 512 parents, pool 14080: 41.4 ms; pool 10240: 41.7; 4 x 256 parents, pool 7040: 41.6; 1 x 1024 parents, pool
 * 28160: 42.2; 2 x 448 on 896 threads (72 registers): 42.3; 2 x 384 on 768 threads (80 registers): 44.2 -- more warps
 * beat more registers.  (With two enumerations per parent and a 8960-entry pool the same geometry took 46.0 ms.) */
using TileDefault = TileConfig<1024, 512, 14080, 1, 2>;
/* small frontiers: shorter tiles, so that every thread group of the grid gets some and the launch lasts a few
 * short tiles instead of one long one (K2 on 2 039 parents: 0.21 -> 0.06 ms) */
using TileSmall = TileConfig<1024, 64, 5120, 1, 4>;
using TileTiny = TileConfig<1024, 32, 2304, 1, 4>;
using TileV1 = TileConfig<1024, 256, 7040, 1, 4>;   /* four 256-parent groups: the default before the inert-move accounting */
using TileV2 = TileConfig<1024, 1024, 28160, 1, 1>; /* one group */
using TileV3 = TileConfig<1024, 512, 10240, 1, 2>;   /* the default with a smaller pool (more L1) */
using TileV4 = TileConfig<896, 448, 15104, 1, 2>;   /* 896 threads: 73 registers */
using TileV5 = TileConfig<768, 384, 16128, 1, 2>;   /* 768 threads: 85 registers, no spills */
constexpr int TILE_VARIANTS = 6;

/* prefix[0..n] = exclusive scan of counts[0..n) */
void scan_counts(pabi_cuda_ctx* ctx, const u32* counts, size_t n, u64* prefix);

/* counts[0..n) of records [first, first+n) of `l`, then prefix[0..n] */
void count_and_scan(pabi_cuda_ctx* ctx, const u64* rec, size_t stride, size_t first, size_t n, u32* counts, u64* prefix) {
  k_count<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, rec, stride, first, n, counts);
  LAUNCHED();
  scan_counts(ctx, counts, n, prefix);
}

void scan_counts(pabi_cuda_ctx* ctx, const u32* counts, size_t n, u64* prefix) {
  const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ctx->tile_sums_cap < tiles) {
    if (ctx->d_tile_sums) CK(cudaFree(ctx->d_tile_sums));
    ctx->d_tile_sums = nullptr;
    CK(cudaMalloc(&ctx->d_tile_sums, (tiles + 1024) * sizeof(u64)));
    ctx->tile_sums_cap = tiles + 1024;
  }
  k_scan_reduce<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(counts, n, ctx->d_tile_sums);
  LAUNCHED();
  k_scan_spine<<<1, 1024, 0, ctx->stream>>>(ctx->d_tile_sums, tiles);
  LAUNCHED();
  k_scan_down<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(counts, n, ctx->d_tile_sums, prefix);
  LAUNCHED();
}

template <TileMode MODE, Accounting ACC, class CFG>
void launch_tile_cfg(pabi_cuda_ctx* ctx, Frontier in, size_t first, size_t n, const u64* prefix, TileOut out) {
  /* the table image alone exceeds the default 48 KB of dynamic shared memory; the attribute is per device, and a
   * context is one device used by one caller at a time, so the record of what has been raised lives in the context */
  const void* fn = (const void*)k_tile<MODE, ACC, CFG>;
  if (!ctx->big_smem.count(fn)) {
    static_assert(sizeof(TileShared<CFG>) <= 232448, "a tile geometry must fit the 227 KB of shared memory a block may use");
    CK(cudaFuncSetAttribute(k_tile<MODE, ACC, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared<CFG>)));
    ctx->big_smem.insert(fn);
  }
  const int grid = grid_for(ctx, n, CFG::PARENTS, CFG::MIN_BLOCKS);
  k_tile<MODE, ACC, CFG><<<grid, CFG::THREADS, sizeof(TileShared<CFG>), ctx->stream>>>(ctx->dt, in, first, n, prefix, out);
  LAUNCHED();
}

template <class CFG>
void launch_leaf_acc(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t n, const TileOut& out) {
  if (acc == PER_ROOT) launch_tile_cfg<TILE_LEAF, PER_ROOT, CFG>(ctx, in, 0, n, nullptr, out);
  else if (acc == WEIGHTED) launch_tile_cfg<TILE_LEAF, WEIGHTED, CFG>(ctx, in, 0, n, nullptr, out);
  else launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, CFG>(ctx, in, 0, n, nullptr, out);
}

/* K2: the last two plies below the n records of `in`.  With n_dev the true size is *n_dev (on the device) and n is the
 * estimate the launch geometry is chosen for. */
void launch_leaf(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t n, const unsigned long long* n_dev = nullptr,
                 const unsigned long long* virt = nullptr) {
  while (ctx->leaf_events.size() < 2 * (ctx->leaf_launches + 1)) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    ctx->leaf_events.push_back(e);
  }
  CK(cudaEventRecord(ctx->leaf_events[2 * ctx->leaf_launches], ctx->stream));
  CK(cudaMemsetAsync(ctx->d_visited + 1, 0, sizeof(unsigned long long), ctx->stream)); /* the tile ticket counter */
  TileOut out{Frontier{nullptr, nullptr, 0, nullptr}, nullptr, nullptr, ctx->d_nodes, ctx->d_visited, ctx->d_visited + 1};
  out.n_in = n_dev, out.virt = virt;
  /* below one default tile per thread group of the grid the frontier is cut into shorter tiles (measured: 197 281
   * parents are still faster with 256-parent tiles, 97 862 with 64-parent ones) */
  const size_t groups = (size_t)ctx->sm_count * TileDefault::GROUPS;
  const size_t small_below = groups * TileDefault::PARENTS, tiny_below = groups * TileSmall::PARENTS;
  if (ctx->tile_variant == 8 && acc == SINGLE_ROOT) { /* warp-synchronous K2 (PABI_TILE_VARIANT=8) */
    k_leaf_warp<SINGLE_ROOT><<<grid_for(ctx, n, 32, 1), LEAFW_THREADS, sizeof(LeafWarpBlock), ctx->stream>>>(ctx->dt, in, n, out);
    LAUNCHED();
  } else if (ctx->tile_variant == 0 && n < small_below) {
    if (n < tiny_below) launch_leaf_acc<TileTiny>(ctx, acc, in, n, out);
    else launch_leaf_acc<TileSmall>(ctx, acc, in, n, out);
  } else if (acc == PER_ROOT) {
    launch_tile_cfg<TILE_LEAF, PER_ROOT, TileDefault>(ctx, in, 0, n, nullptr, out);
  } else if (acc == WEIGHTED) {
    launch_tile_cfg<TILE_LEAF, WEIGHTED, TileDefault>(ctx, in, 0, n, nullptr, out);
  } else {
    switch (ctx->tile_variant) {
      case 1: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileV1>(ctx, in, 0, n, nullptr, out); break;
      case 2: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileV2>(ctx, in, 0, n, nullptr, out); break;
      case 3: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileV3>(ctx, in, 0, n, nullptr, out); break;
      case 4: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileV4>(ctx, in, 0, n, nullptr, out); break;
      case 5: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileV5>(ctx, in, 0, n, nullptr, out); break;
      default: launch_tile_cfg<TILE_LEAF, SINGLE_ROOT, TileDefault>(ctx, in, 0, n, nullptr, out); break;
    }
  }
  CK(cudaEventRecord(ctx->leaf_events[2 * ctx->leaf_launches + 1], ctx->stream));
  ++ctx->leaf_launches;
  if (!n_dev) ctx->leaf_parents += n; /* device-sized frontiers are added when their counters are read back */
}

/* K1 without a sizing pass (TILE_EXPAND_ATOMIC): the children of the records of `in` (n of them, or *n_dev) go to
 * `children` in no particular order; their number is accumulated in *n_out (zeroed at the start of the call). */
template <class CFG>
void launch_expand_atomic_acc(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t n, const TileOut& out) {
  if (acc == PER_ROOT) launch_tile_cfg<TILE_EXPAND_ATOMIC, PER_ROOT, CFG>(ctx, in, 0, n, nullptr, out);
  else launch_tile_cfg<TILE_EXPAND_ATOMIC, SINGLE_ROOT, CFG>(ctx, in, 0, n, nullptr, out);
}

/* The same for a virtual frontier (TILE_EXPAND_VIRTUAL): the children of records [first, first + n) of `in` as 8-byte entries */
template <class CFG>
void launch_expand_virtual_acc(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t first, size_t n, const TileOut& out) {
  if (acc == PER_ROOT) launch_tile_cfg<TILE_EXPAND_VIRTUAL, PER_ROOT, CFG>(ctx, in, first, n, nullptr, out);
  else launch_tile_cfg<TILE_EXPAND_VIRTUAL, SINGLE_ROOT, CFG>(ctx, in, first, n, nullptr, out);
}

void launch_expand_virtual(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t first, size_t n, const unsigned long long* n_dev,
                           unsigned long long* entries, size_t entries_cap, unsigned long long* n_out) {
  TileOut out{Frontier{nullptr, nullptr, 0, nullptr}, nullptr, nullptr, nullptr, nullptr, nullptr};
  out.n_in = n_dev, out.n_out = n_out, out.overflow = ctx->d_fast + FAST_SLOTS - 1, out.out_cap = entries_cap, out.virt_out = entries;
  const size_t groups = (size_t)ctx->sm_count * TileDefault::GROUPS;
  if (n < groups * TileSmall::PARENTS) launch_expand_virtual_acc<TileTiny>(ctx, acc, in, first, n, out);
  else if (n < groups * TileDefault::PARENTS) launch_expand_virtual_acc<TileSmall>(ctx, acc, in, first, n, out);
  else launch_expand_virtual_acc<TileDefault>(ctx, acc, in, first, n, out);
}

void launch_expand_atomic(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t n, const unsigned long long* n_dev, Frontier children,
                          size_t children_cap, unsigned long long* n_out) {
  TileOut out{children, nullptr, nullptr, nullptr, nullptr, nullptr};
  out.n_in = n_dev, out.n_out = n_out, out.overflow = ctx->d_fast + FAST_SLOTS - 1, out.out_cap = children_cap;
  const size_t groups = (size_t)ctx->sm_count * TileDefault::GROUPS;
  if (n < groups * TileSmall::PARENTS) launch_expand_atomic_acc<TileTiny>(ctx, acc, in, n, out);
  else if (n < groups * TileDefault::PARENTS) launch_expand_atomic_acc<TileSmall>(ctx, acc, in, n, out);
  else launch_expand_atomic_acc<TileDefault>(ctx, acc, in, n, out);
}

/* K1: children of records [first, first + n) of `in`, at the offsets of `prefix` (count_and_scan) */
template <class CFG>
void launch_expand_acc(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t first, size_t n, const u64* prefix, const TileOut& out) {
  if (acc == PER_ROOT) launch_tile_cfg<TILE_EXPAND, PER_ROOT, CFG>(ctx, in, first, n, prefix, out);
  else if (acc == WEIGHTED) launch_tile_cfg<TILE_EXPAND, WEIGHTED, CFG>(ctx, in, first, n, prefix, out);
  else launch_tile_cfg<TILE_EXPAND, SINGLE_ROOT, CFG>(ctx, in, first, n, prefix, out);
}

void launch_expand(pabi_cuda_ctx* ctx, Accounting acc, Frontier in, size_t first, size_t n, const u64* prefix, Frontier children) {
  const TileOut out{children, nullptr, nullptr, nullptr, nullptr, nullptr};
  /* small frontiers in shorter tiles, as in launch_leaf: 8 902 parents in 512-parent tiles keep 18 of the 296 thread
   * groups of the grid busy */
  const size_t groups = (size_t)ctx->sm_count * TileDefault::GROUPS;
  if (ctx->tile_variant == 0 && n < groups * TileSmall::PARENTS) launch_expand_acc<TileTiny>(ctx, acc, in, first, n, prefix, out);
  else if (ctx->tile_variant == 0 && n < groups * TileDefault::PARENTS) launch_expand_acc<TileSmall>(ctx, acc, in, first, n, prefix, out);
  else launch_expand_acc<TileDefault>(ctx, acc, in, first, n, prefix, out);
}

/* K3 fused (k_moves_fused): tiles of a batch, its scan words, one launch over the tiles [t0, t1) */
constexpr size_t LIST_TILE = 32;           /* positions per tile: one warp pass */
#ifndef PABI_LIST_CHUNK_TILES
#define PABI_LIST_CHUNK_TILES 2048
#endif
constexpr size_t LIST_CHUNK_TILES = PABI_LIST_CHUNK_TILES; /* 2048 tiles = 65 536 positions = 7.3 MB of records per copy: the copies set the pace (PCIe), so many small chunks keep the tail after the last copy short (measured on 1 M positions, page-locked arrays: 1024 tiles 3.59 ms, 2048 2.66, 4096 2.76, 8192 2.71) */
size_t list_tiles(size_t n) { return (n + LIST_TILE - 1) / LIST_TILE; }

void prepare_list_scan(pabi_cuda_ctx* ctx, size_t tiles) {
  if (ctx->list_status_cap < tiles + 1) {
    if (ctx->d_list_status) CK(cudaFree(ctx->d_list_status));
    ctx->d_list_status = nullptr, ctx->list_status_cap = 0;
    CK(cudaMalloc(&ctx->d_list_status, (tiles + 1025) * sizeof(u64)));
    ctx->list_status_cap = tiles + 1024;
  }
  CK(cudaMemsetAsync(ctx->d_list_status, 0, (tiles + 1) * sizeof(u64), ctx->stream)); /* [tiles] is the ticket counter */
}

void launch_list_chunk(pabi_cuda_ctx* ctx, const ListInput& in, size_t n, size_t t0, size_t t1, u16* moves, size_t moves_cap, u32* offsets) {
  unsigned long long* ticket = ctx->d_list_status + ctx->list_status_cap; /* the last word of the allocation */
  CK(cudaMemsetAsync(ticket, 0, sizeof(u64), ctx->stream));
  const ListScan scan{ctx->d_list_status, ticket, t0, t1};
  const size_t blocks = std::min<size_t>((t1 - t0 + LIST_THREADS / 32 - 1) / (LIST_THREADS / 32), (size_t)ctx->sm_count);
  k_moves_fused<<<(unsigned)blocks, LIST_THREADS, sizeof(ListShared), ctx->stream>>>(ctx->dt, in, n, scan, moves, moves_cap, offsets);
  LAUNCHED();
}

/* K3: ordered move lists of records [0, n) at the offsets of `prefix`; the caller has checked the capacity */
void launch_moves(pabi_cuda_ctx* ctx, const u64* rec, size_t stride, size_t n, const u64* prefix, u16* moves, u32* offsets) {
  if (ctx->tile_variant == 9) { /* the block-level tile pipeline, kept for comparison (PABI_TILE_VARIANT=9) */
    const TileOut out{Frontier{nullptr, nullptr, 0, nullptr}, moves, offsets, nullptr, nullptr, nullptr};
    launch_tile_cfg<TILE_MOVES, SINGLE_ROOT, TileDefault>(ctx, Frontier{const_cast<u64*>(rec), nullptr, stride, nullptr}, 0, n, prefix, out);
    return;
  }
  k_moves_warp<<<grid_for(ctx, n, MOVES_THREADS, 1), MOVES_THREADS, sizeof(MovesShared), ctx->stream>>>(ctx->dt, rec, stride, n, prefix,
                                                                                                      moves, offsets);
  LAUNCHED();
}

/* Walks `remaining` plies below the n records of level `lv`, adding leaf counts
 * to d_nodes (per root when multi). */
/* Transposition merging of the n records of level `lv` (WEIGHTED accounting): identical positions
 * are merged and their multiplicities summed.  Returns the number of survivors, which replace the
 * level's contents (compacted through ctx->merge_level). */
size_t merge_transpositions(pabi_cuda_ctx* ctx, int lv, size_t n) {
  if (n < 2) return n;
  Level& cur = ctx->levels[lv];
  size_t table = 1;
  while (table < 2 * n) table <<= 1;
  if (ctx->merge_table < table) {
    if (ctx->d_tags) {
      CK(cudaFree(ctx->d_tags));
      CK(cudaFree(ctx->d_owner));
    }
    ctx->d_tags = nullptr, ctx->d_owner = nullptr, ctx->merge_table = 0;
    CK(cudaMalloc(&ctx->d_tags, table * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->d_owner, table * sizeof(u32)));
    ctx->merge_table = table;
  }
  u32* slot_of = (u32*)ensure(ctx, S_COUNTS, n * sizeof(u32));
  alloc_level(ctx, ctx->merge_level, cur.cap);
  ensure_mult(ctx, ctx->merge_level);
  CK(cudaMemsetAsync(ctx->d_tags, 0, table * sizeof(unsigned long long), ctx->stream));
  const int grid = grid_for(ctx, n, 256, 8);
  k_dedup_insert<<<grid, 256, 0, ctx->stream>>>(cur.rec, cur.cap, n, ctx->d_tags, table - 1, ctx->d_owner, slot_of, cur.counts);
  LAUNCHED();
  k_dedup_verify<<<grid, 256, 0, ctx->stream>>>(cur.rec, cur.cap, n, ctx->d_owner, slot_of, cur.counts,
                                               (unsigned long long*)cur.mult);
  LAUNCHED();
  scan_counts(ctx, cur.counts, n, cur.prefix); /* counts = survivor flags */
  k_dedup_compact<<<grid, 256, 0, ctx->stream>>>(cur.rec, cur.cap, n, cur.counts, cur.prefix, cur.mult, ctx->merge_level.rec,
                                                ctx->merge_level.cap, ctx->merge_level.mult);
  LAUNCHED();
  CK(cudaMemcpyAsync(ctx->h_result, cur.prefix + n, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::swap(ctx->levels[lv], ctx->merge_level);
  ctx->merged_away += n - (size_t)ctx->h_result[0];
  return (size_t)ctx->h_result[0];
}

/* Walks `remaining` plies below the n records of level `lv`, adding leaf counts
 * to d_nodes (per root for PER_ROOT; times the record's multiplicity for WEIGHTED, where every
 * expanded chunk is passed through merge_transpositions). */
/* Up to max_plies plies below the n records of level lv in ONE launch, as long as the frontier stays within one
 * block (k_root_levels): returns the plies done and leaves the size of the frontier of level lv + plies in n. */
int root_levels(pabi_cuda_ctx* ctx, int lv, size_t& n, int max_plies, Accounting acc) {
  if (acc == WEIGHTED || n == 0 || n > (size_t)ROOT_THREADS || ctx->capacity < (size_t)ROOT_THREADS * 256) return 0;
  if (max_plies > ROOT_MAX_PLIES) max_plies = ROOT_MAX_PLIES;
  if (lv + max_plies >= MAX_LEVELS) max_plies = MAX_LEVELS - 1 - lv;
  if (max_plies <= 0) return 0;
  RootLevels levels;
  levels.level[0] = frontier_of(ctx->levels[lv]);
  for (int k = 1; k <= max_plies; k++) {
    alloc_level(ctx, ctx->levels[lv + k], ctx->capacity);
    levels.level[k] = frontier_of(ctx->levels[lv + k]);
  }
  if (acc == PER_ROOT) k_root_levels<true><<<1, ROOT_THREADS, sizeof(RootShared), ctx->stream>>>(ctx->dt, levels, (u32)n, max_plies, ctx->d_result);
  else k_root_levels<false><<<1, ROOT_THREADS, sizeof(RootShared), ctx->stream>>>(ctx->dt, levels, (u32)n, max_plies, ctx->d_result);
  LAUNCHED();
  CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 3 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  n = (size_t)ctx->h_result[1];
  ctx->interior += ctx->h_result[2];
  return (int)ctx->h_result[0];
}

/* A frontier and what the host knows about its size: `n` records when `n_dev` is null; otherwise the truth is *n_dev, on
 * the device, and n is an estimate that only shapes the launches (every kernel walks its input with a stride or a ticket
 * counter, so any grid covers any size). */
struct Sized {
  size_t n;
  const unsigned long long* n_dev;
};

struct Redo {}; /* a speculative ply did not fit its buffer: the call is repeated on the exact path */

unsigned long long* fast_counter(pabi_cuda_ctx* ctx) { /* a zeroed device counter for one device-sized frontier; null when used up */
  if (ctx->fast_used >= FAST_SLOTS - 1) return nullptr;
  return ctx->d_fast + ctx->fast_used++;
}

/* the exact size of a device-sized frontier: the one place where the host waits in the middle of a walk */
size_t exact_size(pabi_cuda_ctx* ctx, const Sized& s) {
  if (!s.n_dev) return s.n;
  const size_t slot = (size_t)(s.n_dev - ctx->d_fast);
  CK(cudaMemcpyAsync(ctx->h_fast, ctx->d_fast, ctx->fast_used * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream)); /* every counter so far */
  CK(cudaMemcpyAsync(ctx->h_fast + FAST_SLOTS - 1, ctx->d_fast + FAST_SLOTS - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_fast[FAST_SLOTS - 1]) throw Redo{};
  for (auto e = ctx->fast_edges.rbegin(); e != ctx->fast_edges.rend(); ++e) { /* the measured branching of the deepest finished ply */
    const double parents = e->parent_slot >= 0 ? (double)ctx->h_fast[e->parent_slot] : (double)e->parent_n;
    const double children = (double)ctx->h_fast[e->child_slot];
    if (parents > 0 && children > 0) {
      ctx->branching = children / parents;
      break;
    }
  }
  return (size_t)ctx->h_fast[slot];
}

void descend(pabi_cuda_ctx* ctx, int lv, Sized s, int remaining, Accounting acc, bool arrived);

/* Root-subtree sharding (pabi_cuda_perft_shard): when the walk arrives at the split level with a fresh set of records, only
 * those of this call's shard go on (k_filter_shard: a partition by content, independent of the records' order). */
void keep_my_shard(pabi_cuda_ctx* ctx, int lv, Sized& s) {
  alloc_level(ctx, ctx->spare, ctx->capacity);
  unsigned long long* kept = ctx->exact_only ? nullptr : fast_counter(ctx);
  const bool on_device = kept != nullptr;
  if (!kept) {
    kept = (unsigned long long*)ctx->d_result + 3;
    CK(cudaMemsetAsync(kept, 0, sizeof(u64), ctx->stream));
  }
  k_filter_shard<<<grid_for(ctx, s.n, 256, 8), 256, 0, ctx->stream>>>(frontier_of(ctx->levels[lv]), s.n, s.n_dev, ctx->shard_index,
                                                                      ctx->shard_count, frontier_of(ctx->spare), kept);
  LAUNCHED();
  std::swap(ctx->levels[lv], ctx->spare);
  if (on_device) {
    s = Sized{s.n / ctx->shard_count + 1, kept};
  } else {
    CK(cudaMemcpyAsync(ctx->h_result + 3, kept, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s = Sized{(size_t)ctx->h_result[3], nullptr};
  }
}

void count_level(pabi_cuda_ctx* ctx, int lv, const Sized& s, Accounting acc) { /* one ply left: position.rs:914-916 */
  Level& cur = ctx->levels[lv];
  const int grid = grid_for(ctx, s.n, FLAT_THREADS, 4);
  if (acc == PER_ROOT)
    k_count_accumulate<PER_ROOT><<<grid, FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, frontier_of(cur), s.n, s.n_dev, ctx->d_nodes);
  else if (acc == WEIGHTED)
    k_count_accumulate<WEIGHTED><<<grid, FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, frontier_of(cur), s.n, s.n_dev, ctx->d_nodes);
  else
    k_count_accumulate<SINGLE_ROOT><<<grid, FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, frontier_of(cur), s.n, s.n_dev, ctx->d_nodes);
  LAUNCHED();
}

void note_size(pabi_cuda_ctx* ctx, const Sized& s, bool leaf_input) { /* statistics of a frontier that is being walked */
  if (!s.n_dev) {
    ctx->interior += s.n;
    return;
  }
  const int slot = (int)(s.n_dev - ctx->d_fast);
  ctx->fast_interior.push_back(slot);
  if (leaf_input) ctx->fast_leaf.push_back(slot);
}

/* The exact path: count -> scan -> expand with the children in the reference's move order; a level whose children would
 * not fit the next buffer is cut into chunks, each finished depth-first before the next starts. */
void descend_exact(pabi_cuda_ctx* ctx, int lv, size_t n, int remaining, Accounting acc) {
  if (n == 0) return;
  if (remaining > 2 && ctx->exact_only) { /* the launch-bound top of the tree: several plies per launch, two are left for K2 */
    size_t below = n;
    const int plies = root_levels(ctx, lv, below, std::min(remaining - 2, ctx->shard_level > lv ? ctx->shard_level - lv : remaining), acc);
    if (plies) {
      descend(ctx, lv + plies, Sized{below, nullptr}, remaining - plies, acc, true);
      return;
    }
  }
  if (lv + 1 >= MAX_LEVELS) {
    ctx->error = "perft depth exceeds the supported number of breadth-first levels";
    throw CudaFail{cudaErrorInvalidValue};
  }
  ctx->interior += n;
  alloc_level(ctx, ctx->levels[lv + 1], ctx->capacity);
  if (acc == WEIGHTED) ensure_mult(ctx, ctx->levels[lv + 1]);
  {
    Level& cur = ctx->levels[lv];
    count_and_scan(ctx, cur.rec, cur.cap, 0, n, cur.counts, cur.prefix);
  }
  size_t start = 0;
  while (start < n) {
    /* levels may have been swapped by a deeper merge_transpositions or shard filter: re-read them every time */
    Level& cur = ctx->levels[lv];
    Level& next = ctx->levels[lv + 1];
    size_t end = n, children = 0;
    bool whole = false;
    if (start == 0) { /* the usual case, all children fit the next level: the total is the last prefix entry, no search */
      CK(cudaMemcpyAsync(ctx->h_result, cur.prefix + n, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      children = (size_t)ctx->h_result[0];
      whole = children <= next.cap;
    }
    if (!whole) {
      k_chunk_end<<<1, 1, 0, ctx->stream>>>(cur.prefix, start, n, (u64)next.cap, ctx->d_result);
      LAUNCHED();
      CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      end = (size_t)ctx->h_result[0];
      children = (size_t)ctx->h_result[1];
    }
    if (end == start) {
      ctx->error = "frontier capacity is smaller than one position's move list";
      throw CudaFail{cudaErrorInvalidValue};
    }
    if (children) {
      launch_expand(ctx, acc, frontier_of(cur), start, end - start, cur.prefix, frontier_of(next));
      if (acc == WEIGHTED) children = merge_transpositions(ctx, lv + 1, children);
      descend(ctx, lv + 1, Sized{children, nullptr}, remaining - 1, acc, true);
    }
    start = end;
  }
}

/* Three plies left: the ply below level lv is not materialised.  K1 writes every child of the level as an 8-byte
 * (parent index, kind, move) entry (TILE_EXPAND_VIRTUAL, into the next level's record buffer, which holds 8 entries per
 * record slot) and K2 makes each of its parents from such an entry and the parent's record, then walks the last two plies as
 * always.  Against the materialised ply this writes and reads 8 instead of 64 bytes per position and needs no sizing pass;
 * the price is one make_move per K2 parent inside the ALU-bound K2.  The level's records go in chunks sized from the
 * measured (or assumed) branching; a chunk whose entries overflow the buffer after all raises the flag (-> exact redo). */
bool virtual_ply(pabi_cuda_ctx* ctx, int lv, Sized s, Accounting acc) {
  if (ctx->capacity >= (size_t(1) << 32)) return false; /* an entry holds a 32-bit record index */
  alloc_level(ctx, ctx->levels[lv + 1], ctx->capacity);
  const size_t entries_cap = ctx->levels[lv + 1].cap * POS_WORDS;
  unsigned long long* entries = (unsigned long long*)ctx->levels[lv + 1].rec;
  if (s.n_dev && (double)s.n * (double)(BRANCH_ESTIMATE * BRANCH_MARGIN) <= (double)entries_cap) {
    /* a small level sized on the device: one chunk is certain to do, so the host need not learn the size */
    unsigned long long* made = fast_counter(ctx);
    if (!made) return false;
    note_size(ctx, s, false);
    launch_expand_virtual(ctx, acc, frontier_of(ctx->levels[lv]), 0, s.n, s.n_dev, entries, entries_cap, made);
    const int slot = (int)(made - ctx->d_fast);
    ctx->fast_interior.push_back(slot), ctx->fast_leaf.push_back(slot);
    launch_leaf(ctx, acc, frontier_of(ctx->levels[lv]), s.n * BRANCH_ESTIMATE, made, entries);
    return true;
  }
  const size_t n = exact_size(ctx, s); /* chunk boundaries are host decisions */
  if (n == 0) return true;
  const double per_parent = ctx->branching > 0 ? ctx->branching * 1.6 : (double)(BRANCH_ESTIMATE * BRANCH_MARGIN);
  const size_t chunk = (size_t)std::min<double>((double)n, (double)entries_cap / per_parent);
  if (chunk == 0 || (n + chunk - 1) / chunk > (size_t)(FAST_SLOTS - 1 - ctx->fast_used)) return false;
  ctx->interior += n;
  for (size_t start = 0; start < n; start += chunk) {
    const size_t len = std::min(chunk, n - start);
    unsigned long long* made = fast_counter(ctx);
    launch_expand_virtual(ctx, acc, frontier_of(ctx->levels[lv]), start, len, nullptr, entries, entries_cap, made);
    const int slot = (int)(made - ctx->d_fast);
    ctx->fast_interior.push_back(slot), ctx->fast_leaf.push_back(slot);
    launch_leaf(ctx, acc, frontier_of(ctx->levels[lv]), (size_t)((double)len * per_parent / 1.6), made, entries);
  }
  return true;
}

/* Walks `remaining` plies below the records of level `lv`, adding leaf counts to d_nodes (per root for PER_ROOT; times the
 * record's multiplicity for WEIGHTED, where every expanded chunk is passed through merge_transpositions).
 *
 * Small plies are expanded WITHOUT the host learning their size (TILE_EXPAND_ATOMIC: the next frontier's size is a counter on
 * the device, the following launch reads it): the launch-bound part of a walk runs without a single host round trip.  A ply
 * is taken this way only while four times the estimated number of children still fits the next buffer; when the estimate
 * gets large the host fetches the exact size (exact_size) and continues on the exact, chunking path, where a round trip is
 * negligible.  Should a speculative ply overflow all the same, its kernel raises a flag and the call is redone exactly. */
void descend(pabi_cuda_ctx* ctx, int lv, Sized s, int remaining, Accounting acc, bool arrived) {
  if (!s.n_dev && s.n == 0) return;
  if (arrived && lv == ctx->shard_level && ctx->shard_count > 1) {
    keep_my_shard(ctx, lv, s);
    if (!s.n_dev && s.n == 0) return;
  }
  if (remaining == 1) {
    note_size(ctx, s, false);
    count_level(ctx, lv, s, acc);
    return;
  }
  const bool split_below = ctx->shard_count > 1 && ctx->shard_level > lv; /* the shard is cut at a deeper ply: K2 must not pass over it */
  if (remaining == 2 && !split_below) { /* the children are generated and counted inside K2 */
    note_size(ctx, s, true);
    launch_leaf(ctx, acc, frontier_of(ctx->levels[lv]), s.n, s.n_dev);
    return;
  }
  const bool may_speculate = !ctx->exact_only && acc != WEIGHTED && lv + 1 < MAX_LEVELS;
  if (remaining == 3 && may_speculate && !split_below && !ctx->no_virtual && virtual_ply(ctx, lv, s, acc)) return;
  unsigned long long* n_children = nullptr;
  bool speculate = may_speculate && s.n * BRANCH_ESTIMATE * BRANCH_MARGIN <= ctx->capacity && (n_children = fast_counter(ctx)) != nullptr;
  if (!speculate) {
    /* the estimate is too large for the buffer: fetch the exact size.  If the branching of the plies above is known by
     * then (it is whenever they were speculative), one more ply may still go without its sizing pass: this size x the
     * measured branching x 1.6 must fit (sides alternate, so consecutive plies differ; a miss costs a redo, never a count) */
    const size_t n = exact_size(ctx, s);
    if (n == 0) return;
    s = Sized{n, nullptr};
    speculate = may_speculate && s.n_dev == nullptr && ctx->branching > 0 &&
                (double)n * ctx->branching * 1.6 <= (double)ctx->capacity && (n_children = fast_counter(ctx)) != nullptr;
    if (!speculate) {
      descend_exact(ctx, lv, n, remaining, acc);
      return;
    }
  }
  alloc_level(ctx, ctx->levels[lv + 1], ctx->capacity);
  note_size(ctx, s, false);
  launch_expand_atomic(ctx, acc, frontier_of(ctx->levels[lv]), s.n, s.n_dev, frontier_of(ctx->levels[lv + 1]), ctx->levels[lv + 1].cap,
                       n_children);
  ctx->fast_edges.push_back({s.n, s.n_dev ? (int)(s.n_dev - ctx->d_fast) : -1, (int)(n_children - ctx->d_fast)});
  const size_t estimate = ctx->branching > 0 && !s.n_dev ? (size_t)((double)s.n * ctx->branching) : s.n * BRANCH_ESTIMATE;
  descend(ctx, lv + 1, Sized{estimate, n_children}, remaining - 1, acc, true);
}

struct CallTimer {
  pabi_cuda_ctx* ctx;
  std::chrono::steady_clock::time_point t0;
  explicit CallTimer(pabi_cuda_ctx* c) : ctx(c), t0(std::chrono::steady_clock::now()) {
    ctx->launches = 0, ctx->interior = 0, ctx->leaf_launches = 0, ctx->leaf_parents = 0;
    ctx->error.clear();
  }
  void begin_device() { CK(cudaEventRecord(ctx->ev0, ctx->stream)); }
  void end_device() { CK(cudaEventRecord(ctx->ev1, ctx->stream)); }
  void finish(pabi_timing_t* t, uint64_t bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_bad_roots) { /* set by upload_roots' k_pack in this call */
      const unsigned bad = *ctx->h_bad_roots;
      *ctx->h_bad_roots = 0;
      ctx->error = std::to_string(bad) + " input record(s) are not positions (each side needs exactly one king, no pawns on the "
                   "back ranks): build positions with pabi_position_from_fen / pabi_position_parse";
      throw CudaFail{(cudaError_t)-5};
    }
    if (!t) return;
    std::memset(t, 0, sizeof *t);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    t->device_ms = ms;
    t->host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    t->algorithmic_bytes = bytes;
    t->kernel_launches = ctx->launches;
    t->interior_nodes = ctx->interior;
    for (uint64_t i = 0; i < ctx->leaf_launches; i++) {
      CK(cudaEventElapsedTime(&ms, ctx->leaf_events[2 * i], ctx->leaf_events[2 * i + 1]));
      t->leaf_kernel_ms += ms;
    }
    t->leaf_kernel_launches = ctx->leaf_launches;
    t->leaf_parents = ctx->leaf_parents;
  }
};

/* Upload n positions (host AoS) into level 0 as SoA with roots = index. */
void upload_roots(pabi_cuda_ctx* ctx, const pabi_position_t* pos, size_t n) {
  alloc_level(ctx, ctx->levels[0], n > 4096 ? n : 4096);
  void* staging = ensure(ctx, S_STAGING, n * sizeof(pabi_position_t));
  CK(cudaMemcpyAsync(staging, pos, n * sizeof(pabi_position_t), cudaMemcpyHostToDevice, ctx->stream));
  Level& l0 = ctx->levels[0];
  CK(cudaMemsetAsync(ctx->d_bad_roots, 0, sizeof(u32), ctx->stream));
  k_pack<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>((const PositionImage*)staging, n, l0.rec, l0.cap, 0, l0.roots, ctx->d_bad_roots);
  LAUNCHED();
  CK(cudaMemcpyAsync(ctx->h_bad_roots, ctx->d_bad_roots, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream)); /* read in CallTimer::finish */
}

int fail(pabi_cuda_ctx* ctx, const CudaFail& f) {
  if ((int)f.code < 0) return (int)f.code; /* an argument error found on the device; the message is set */
  if (ctx->error.empty()) ctx->error = cudaGetErrorString(f.code);
  cudaGetLastError();
  return (int)f.code;
}

int bad_arg(pabi_cuda_ctx* ctx, const char* what) {
  if (ctx) ctx->error = what;
  return -1;
}

int perft_impl(pabi_cuda_ctx* ctx, const pabi_position_t* roots, size_t n, int depth, uint64_t* nodes_out, Accounting acc,
               uint32_t split_ply, uint32_t shard_index, uint32_t shard_count, pabi_timing_t* timing) {
  const bool multi = acc == PER_ROOT;
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (depth == 0) {
      for (size_t i = 0; i < n; i++) nodes_out[i] = (shard_index == 0) ? 1 : 0;
      timer.begin_device(), timer.end_device();
      timer.finish(timing, 0);
      return 0;
    }
    const size_t n_out = multi ? n : 1;
    ensure_nodes(ctx, n_out);
    upload_roots(ctx, roots, n);
    ctx->merged_away = 0;
    if (acc == WEIGHTED) { /* the root is reached by one path */
      ensure_mult(ctx, ctx->levels[0]);
      const u64 one = 1;
      CK(cudaMemcpyAsync(ctx->levels[0].mult, &one, sizeof one, cudaMemcpyHostToDevice, ctx->stream));
    }
    timer.begin_device();
    /* root-subtree sharding: the walk keeps one shard of the frontier at the split ply (at most depth - 1, so that a ply is left) */
    ctx->shard_index = shard_index, ctx->shard_count = shard_count ? shard_count : 1;
    ctx->shard_level = shard_count > 1 ? (int)std::min<uint32_t>(split_ply, (uint32_t)depth - 1) : -1;
    for (int attempt = 0;; attempt++) {
      ctx->exact_only = ctx->force_exact || attempt > 0;
      ctx->fast_used = 0, ctx->fast_interior.clear(), ctx->fast_leaf.clear(), ctx->fast_edges.clear(), ctx->branching = 0;
      ctx->interior = 0, ctx->leaf_parents = 0, ctx->leaf_launches = 0;
      CK(cudaMemsetAsync(ctx->d_nodes, 0, n_out * sizeof(unsigned long long), ctx->stream));
      CK(cudaMemsetAsync(ctx->d_visited, 0, sizeof(unsigned long long), ctx->stream));
      if (!ctx->exact_only) CK(cudaMemsetAsync(ctx->d_fast, 0, FAST_SLOTS * sizeof(u64), ctx->stream));
      bool redo = false;
      try {
        descend(ctx, 0, Sized{n, nullptr}, depth, acc, true);
      } catch (const Redo&) {
        redo = true;
      }
      if (!redo) {
        timer.end_device();
        CK(cudaMemcpyAsync(nodes_out, ctx->d_nodes, n_out * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_result + 2, ctx->d_visited, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->fast_used) { /* the sizes of the device-sized frontiers and the overflow flag, in the same round trip */
          CK(cudaMemcpyAsync(ctx->h_fast, ctx->d_fast, ctx->fast_used * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
          CK(cudaMemcpyAsync(ctx->h_fast + FAST_SLOTS - 1, ctx->d_fast + FAST_SLOTS - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        redo = ctx->fast_used && ctx->h_fast[FAST_SLOTS - 1] != 0;
      }
      if (!redo) break;
      if (attempt > 0) {
        ctx->error = "the exact path reported an overflow";
        throw CudaFail{cudaErrorUnknown};
      }
      if (acc == WEIGHTED) ctx->merged_away = 0;
    }
#ifdef PABI_STATS
    {
      unsigned long long st[3];
      CK(cudaMemcpy(st, ctx->d_visited + 2, sizeof st, cudaMemcpyDeviceToHost));
      std::fprintf(stderr, "K2 stats: %llu pool passes, %llu overflowed passes, fullest pool %llu entries\n", st[0], st[1], st[2]);
      CK(cudaMemset(ctx->d_visited + 2, 0, sizeof st));
    }
#endif
    for (int slot : ctx->fast_interior) ctx->interior += ctx->h_fast[slot];
    for (int slot : ctx->fast_leaf) ctx->leaf_parents += ctx->h_fast[slot];
    ctx->interior += ctx->h_result[2];
    /* every node of plies 1..d-1 written once and read once at 64 B */
    timer.finish(timing, 2 * 64 * (ctx->interior > n ? ctx->interior - n : 0));
    if (timing) timing->leaf_children = ctx->h_result[2], timing->merged_records = ctx->merged_away;
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  } catch (const std::bad_alloc&) {
    ctx->error = "out of host memory";
    return -3;
  }
}

} /* namespace */

extern "C" {

int pabi_cuda_create(int device, size_t frontier_capacity, pabi_cuda_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  pabi_cuda_ctx* ctx = new (std::nothrow) pabi_cuda_ctx;
  if (!ctx) return -3;
  try {
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    ctx->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (frontier_capacity) ctx->capacity = frontier_capacity < 4096 ? 4096 : frontier_capacity;
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ctx->ev0));
    CK(cudaEventCreate(&ctx->ev1));
    const HostTables& h = host_tables();
    CK(cudaMalloc(&ctx->d_tables, sizeof(SharedTables)));
    CK(cudaMalloc(&ctx->d_rook, sizeof h.rook_attacks));
    CK(cudaMalloc(&ctx->d_rays, sizeof h.rays));
    CK(cudaMemcpy(ctx->d_tables, &h.shared, sizeof(SharedTables), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_rook, h.rook_attacks, sizeof h.rook_attacks, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_rays, h.rays, sizeof h.rays, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->d_zobrist, sizeof h.zobrist));
    CK(cudaMemcpy(ctx->d_zobrist, h.zobrist, sizeof h.zobrist, cudaMemcpyHostToDevice));
    ctx->dt.image = (const SharedTables*)ctx->d_tables;
    ctx->dt.g.rook_attacks = ctx->d_rook;
    ctx->dt.g.rays = ctx->d_rays;
    ctx->dt.g.zobrist = ctx->d_zobrist;
    CK(cudaMalloc(&ctx->d_fast, FAST_SLOTS * sizeof(u64)));
    CK(cudaMallocHost(&ctx->h_fast, FAST_SLOTS * sizeof(u64)));
    if (const char* v = std::getenv("PABI_EXACT_LEVELS")) ctx->force_exact = std::atoi(v) != 0; /* no speculative plies (comparison runs) */
    if (const char* v = std::getenv("PABI_NO_VIRTUAL_PLY")) ctx->no_virtual = std::atoi(v) != 0;
    CK(cudaMalloc(&ctx->d_bad_roots, sizeof(u32)));
    CK(cudaMallocHost(&ctx->h_bad_roots, sizeof(u32)));
    *ctx->h_bad_roots = 0;
    CK(cudaMalloc(&ctx->d_result, 4 * sizeof(u64)));
    CK(cudaMalloc(&ctx->d_visited, 8 * sizeof(unsigned long long))); /* [0] statistics, [1] tile tickets, [2..4] PABI_STATS builds */
    CK(cudaMemset(ctx->d_visited, 0, 8 * sizeof(unsigned long long)));
    CK(cudaMallocHost(&ctx->h_result, 4 * sizeof(u64)));
    /* kernels that stage the table image need more than the default 48 KB of dynamic shared memory */
    CK(cudaFuncSetAttribute(k_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_root_levels<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RootShared)));
    CK(cudaFuncSetAttribute(k_root_levels<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RootShared)));
    CK(cudaFuncSetAttribute(k_count_accumulate<SINGLE_ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_count_accumulate<PER_ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_count_accumulate<WEIGHTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_in_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_attack_info, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_make_moves, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_child_hashes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_random_playouts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_leaf_warp<SINGLE_ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeafWarpBlock)));
    CK(cudaFuncSetAttribute(k_moves_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MovesShared)));
    CK(cudaFuncSetAttribute(k_moves_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ListShared)));
    CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    if (const char* v = std::getenv("PABI_LIST_TWO_PASS")) ctx->list_two_pass = std::atoi(v) != 0;
    CK(cudaFuncSetAttribute(k_parse_fens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_games_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_games_result, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    CK(cudaFuncSetAttribute(k_games_play_random, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM));
    if (const char* v = std::getenv("PABI_TILE_VARIANT")) {
      ctx->tile_variant = std::atoi(v);
      if (ctx->tile_variant != 9 && ctx->tile_variant != 8 && (ctx->tile_variant < 0 || ctx->tile_variant >= TILE_VARIANTS)) ctx->tile_variant = 0;
      if (ctx->tile_variant == 9) ctx->list_two_pass = true; /* the block-level K3 needs the scanned offsets */
      if (ctx->tile_variant == 8) ctx->no_virtual = true;    /* the warp-synchronous K2 reads materialised parents only */
    }
  } catch (const CudaFail& f) {
    std::fprintf(stderr, "pabi_cuda_create: %s\n", ctx->error.c_str());
    int code = (int)f.code;
    pabi_cuda_destroy(ctx);
    return code;
  }
  *out = ctx;
  return 0;
}

void pabi_cuda_destroy(pabi_cuda_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  auto free_level = [](Level& l) {
    if (l.rec) cudaFree(l.rec), cudaFree(l.roots), cudaFree(l.counts), cudaFree(l.prefix), cudaFree(l.mult);
    l = Level();
  };
  for (Level& l : ctx->levels) free_level(l);
  free_level(ctx->spare);
  free_level(ctx->merge_level);
  if (ctx->d_tags) cudaFree(ctx->d_tags), cudaFree(ctx->d_owner);
  for (Buffer& b : ctx->scratch)
    if (b.ptr) cudaFree(b.ptr);
  if (ctx->d_tile_sums) cudaFree(ctx->d_tile_sums);
  if (ctx->d_result) cudaFree(ctx->d_result);
  if (ctx->d_visited) cudaFree(ctx->d_visited);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->d_nodes) cudaFree(ctx->d_nodes);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_rook) cudaFree(ctx->d_rook);
  if (ctx->d_rays) cudaFree(ctx->d_rays);
  if (ctx->d_zobrist) cudaFree(ctx->d_zobrist);
  if (ctx->d_bad_roots) cudaFree(ctx->d_bad_roots);
  if (ctx->d_fast) cudaFree(ctx->d_fast);
  if (ctx->h_fast) cudaFreeHost(ctx->h_fast);
  if (ctx->h_bad_roots) cudaFreeHost(ctx->h_bad_roots);
  for (cudaEvent_t e : ctx->leaf_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->list_events) cudaEventDestroy(e);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  if (ctx->d_list_status) cudaFree(ctx->d_list_status);
  if (ctx->h_list_totals) cudaFreeHost(ctx->h_list_totals);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* pabi_cuda_last_error(pabi_cuda_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
int pabi_cuda_device(const pabi_cuda_ctx* ctx) { return ctx ? ctx->device : -1; }

int pabi_cuda_bounds_violation(pabi_cuda_ctx* ctx) {
#ifdef PABI_BOUNDS
  if (!ctx) return -1;
  unsigned int line = 0, zero = 0;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
      cudaMemcpyFromSymbol(&line, g_bounds_violation, sizeof line) != cudaSuccess ||
      cudaMemcpyToSymbol(g_bounds_violation, &zero, sizeof zero) != cudaSuccess)
    return -2;
  return (int)line;
#else
  (void)ctx;
  return -1;
#endif
}

int pabi_cuda_host_alloc(size_t bytes, void** out) {
  if (!out) return -1;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) cudaGetLastError();
  return (int)e;
}

void pabi_cuda_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int pabi_cuda_perft(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, uint64_t* nodes, pabi_timing_t* timing) {
  if (!ctx || !root || !nodes) return bad_arg(ctx, "null argument");
  return perft_impl(ctx, root, 1, depth, nodes, SINGLE_ROOT, 0, 0, 1, timing);
}

int pabi_cuda_perft_hashed(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, uint64_t* nodes, pabi_timing_t* timing) {
  if (!ctx || !root || !nodes) return bad_arg(ctx, "null argument");
  return perft_impl(ctx, root, 1, depth, nodes, WEIGHTED, 0, 0, 1, timing);
}

int pabi_cuda_perft_batch(pabi_cuda_ctx* ctx, const pabi_position_t* roots, size_t n, uint8_t depth, uint64_t* nodes_out,
                          pabi_timing_t* timing) {
  if (!ctx || (n && (!roots || !nodes_out))) return bad_arg(ctx, "null argument");
  if (n == 0) {
    if (timing) std::memset(timing, 0, sizeof *timing);
    return 0;
  }
  if (n > ctx->capacity) return bad_arg(ctx, "more roots than the frontier capacity");
  return perft_impl(ctx, roots, n, depth, nodes_out, PER_ROOT, 0, 0, 1, timing);
}

int pabi_cuda_perft_batch_depths(pabi_cuda_ctx* ctx, const pabi_position_t* roots, const uint8_t* depths, size_t n,
                                 uint64_t* nodes_out, pabi_timing_t* timing) {
  if (!ctx || (n && (!roots || !depths || !nodes_out))) return bad_arg(ctx, "null argument");
  if (timing) std::memset(timing, 0, sizeof *timing);
  if (n > ctx->capacity) return bad_arg(ctx, "more roots than the frontier capacity");
  /* roots of equal depth walk together: one per-root pass per distinct depth */
  std::map<uint8_t, std::vector<size_t>> groups;
  for (size_t i = 0; i < n; i++) groups[depths[i]].push_back(i);
  std::vector<pabi_position_t> batch;
  std::vector<uint64_t> counts;
  for (const auto& [depth, members] : groups) {
    batch.clear();
    for (size_t i : members) batch.push_back(roots[i]);
    counts.assign(members.size(), 0);
    pabi_timing_t t;
    const int rc = perft_impl(ctx, batch.data(), batch.size(), depth, counts.data(), PER_ROOT, 0, 0, 1, &t);
    if (rc) return rc;
    for (size_t k = 0; k < members.size(); k++) nodes_out[members[k]] = counts[k];
    if (timing) {
      timing->device_ms += t.device_ms, timing->host_ms += t.host_ms, timing->algorithmic_bytes += t.algorithmic_bytes;
      timing->kernel_launches += t.kernel_launches, timing->interior_nodes += t.interior_nodes;
      timing->leaf_kernel_ms += t.leaf_kernel_ms, timing->leaf_kernel_launches += t.leaf_kernel_launches;
      timing->leaf_parents += t.leaf_parents, timing->leaf_children += t.leaf_children;
    }
  }
  return 0;
}

int pabi_cuda_perft_shard(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, uint32_t split_ply,
                          uint32_t shard_index, uint32_t shard_count, uint64_t* nodes, pabi_timing_t* timing) {
  if (!ctx || !root || !nodes) return bad_arg(ctx, "null argument");
  if (shard_count == 0 || shard_index >= shard_count) return bad_arg(ctx, "shard_index must be < shard_count");
  return perft_impl(ctx, root, 1, depth, nodes, SINGLE_ROOT, split_ply, shard_index, shard_count, timing);
}

int pabi_cuda_perft_multi(pabi_cuda_ctx* const* ctxs, size_t n_ctx, const pabi_position_t* root, uint8_t depth,
                          uint32_t split_ply, uint64_t* nodes, double* per_ctx_ms) {
  if (!ctxs || n_ctx == 0 || !root || !nodes) return -1;
  for (size_t i = 0; i < n_ctx; i++)
    if (!ctxs[i]) return -1;
  std::vector<uint64_t> part(n_ctx, 0);
  std::vector<int> rc(n_ctx, 0);
  std::vector<pabi_timing_t> timing(n_ctx);
  std::vector<std::thread> workers;
  for (size_t i = 0; i < n_ctx; i++)
    workers.emplace_back([&, i] {
      rc[i] = pabi_cuda_perft_shard(ctxs[i], root, depth, split_ply, (uint32_t)i, (uint32_t)n_ctx, &part[i], &timing[i]);
    });
  for (std::thread& w : workers) w.join();
  uint64_t total = 0;
  for (size_t i = 0; i < n_ctx; i++) {
    if (rc[i]) return rc[i];
    total += part[i];
    if (per_ctx_ms) per_ctx_ms[i] = timing[i].device_ms;
  }
  *nodes = total;
  return 0;
}

int pabi_cuda_count_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint32_t* counts,
                                pabi_timing_t* timing) {
  if (!ctx || (n && (!positions || !counts))) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    upload_roots(ctx, positions, n);
    Level& l0 = ctx->levels[0];
    timer.begin_device();
    k_count<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, l0.rec, l0.cap, 0, n, l0.counts);
    LAUNCHED();
    ctx->interior += n;
    timer.end_device();
    CK(cudaMemcpyAsync(counts, l0.counts, n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 68 * n);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_in_check_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint8_t* in_check,
                             pabi_timing_t* timing) {
  if (!ctx || (n && (!positions || !in_check))) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    upload_roots(ctx, positions, n);
    Level& l0 = ctx->levels[0];
    u8* d_out = (u8*)ensure(ctx, S_SMALL_A, n);
    timer.begin_device();
    k_in_check<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, l0.rec, l0.cap, n, d_out);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(in_check, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 65 * n);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_attack_info_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint64_t* out,
                                pabi_timing_t* timing) {
  if (!ctx || (n && (!positions || !out))) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    upload_roots(ctx, positions, n);
    Level& l0 = ctx->levels[0];
    u64* d_out = (u64*)ensure(ctx, S_SMALL_A, 5 * n * sizeof(u64));
    timer.begin_device();
    k_attack_info<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, l0.rec, l0.cap, n, d_out);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(out, d_out, 5 * n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 104 * n);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_random_playouts(pabi_cuda_ctx* ctx, const pabi_position_t* starts, const uint64_t* seeds, size_t n_games,
                              uint32_t max_plies, uint32_t stride, uint32_t per_game_cap, pabi_position_t* out,
                              uint32_t* counts, pabi_timing_t* timing) {
  if (!ctx || (n_games && (!starts || !seeds || !out || !counts))) return bad_arg(ctx, "null argument");
  if (stride == 0 || per_game_cap == 0) return bad_arg(ctx, "stride and per_game_cap must be positive");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n_games == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    upload_roots(ctx, starts, n_games);
    Level& l0 = ctx->levels[0];
    const size_t slots = n_games * (size_t)per_game_cap;
    u64* d_seeds = (u64*)ensure(ctx, S_SMALL_A, n_games * sizeof(u64));
    u32* d_counts = (u32*)ensure(ctx, S_SMALL_B, n_games * sizeof(u32));
    PositionImage* d_out = (PositionImage*)ensure(ctx, S_AOS_OUT, slots * sizeof(PositionImage));
    CK(cudaMemcpyAsync(d_seeds, seeds, n_games * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    timer.begin_device();
    k_random_playouts<<<grid_for(ctx, n_games, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(
        ctx->dt, l0.rec, l0.cap, n_games, (const PositionImage*)ctx->scratch[S_STAGING].ptr, d_seeds, max_plies, stride, per_game_cap,
        d_out, d_counts);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(counts, d_counts, n_games * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out, d_out, slots * sizeof(PositionImage), cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 0);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

} /* extern "C" */

struct pabi_cuda_games {
  pabi_cuda_ctx* ctx = nullptr;
  GamesState gs{};
  u16* d_actions = nullptr;
  u8* d_results = nullptr;
  u32* d_applied = nullptr;
  u64* d_seeds = nullptr;
};

namespace {
int games_overflowed(pabi_cuda_games* games) {
  pabi_cuda_ctx* ctx = games->ctx;
  u32 overflow = 0;
  CK(cudaMemcpyAsync(&overflow, games->gs.overflow, sizeof overflow, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (overflow) {
    CK(cudaMemsetAsync(games->gs.overflow, 0, sizeof(u32), ctx->stream));
    ctx->error = "a game received more actions than history_cap";
    return -4;
  }
  return 0;
}
} /* namespace */

extern "C" {

int pabi_cuda_games_create(pabi_cuda_ctx* ctx, const pabi_position_t* roots, size_t n, uint32_t history_cap,
                           pabi_cuda_games** out) {
  if (!ctx || !out || !roots || n == 0) return bad_arg(ctx, "null argument or no games");
  *out = nullptr;
  pabi_cuda_games* games = new (std::nothrow) pabi_cuda_games;
  if (!games) return -3;
  games->ctx = ctx;
  try {
    CK(cudaSetDevice(ctx->device));
    GamesState& gs = games->gs;
    gs.n = n, gs.history_cap = history_cap;
    CK(cudaMalloc(&gs.rec, n * POS_WORDS * sizeof(u64)));
    CK(cudaMalloc(&gs.keys, n * ((size_t)history_cap + 1) * sizeof(u64)));
    CK(cudaMalloc(&gs.plies, n * sizeof(u32)));
    CK(cudaMalloc(&gs.perspective, n));
    CK(cudaMalloc(&gs.threefold, n));
    CK(cudaMalloc(&gs.overflow, sizeof(u32)));
    CK(cudaMalloc(&games->d_actions, n * sizeof(u16)));
    CK(cudaMalloc(&games->d_results, n));
    CK(cudaMalloc(&games->d_applied, n * sizeof(u32)));
    CK(cudaMalloc(&games->d_seeds, n * sizeof(u64)));
    CK(cudaMemsetAsync(gs.overflow, 0, sizeof(u32), ctx->stream));
    void* staging = ensure(ctx, S_STAGING, n * sizeof(pabi_position_t));
    CK(cudaMemcpyAsync(staging, roots, n * sizeof(pabi_position_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bad_roots, 0, sizeof(u32), ctx->stream));
    k_pack<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>((const PositionImage*)staging, n, gs.rec, n, 0, nullptr, ctx->d_bad_roots);
    CK(cudaGetLastError());
    k_games_init<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(gs, (const PositionImage*)staging);
    CK(cudaGetLastError());
    u32 bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->d_bad_roots, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) {
      ctx->error = "a root is not a position (each side needs exactly one king, no pawns on the back ranks)";
      throw CudaFail{(cudaError_t)-5};
    }
  } catch (const CudaFail& f) {
    pabi_cuda_games_destroy(games);
    return fail(ctx, f);
  }
  *out = games;
  return 0;
}

void pabi_cuda_games_destroy(pabi_cuda_games* games) {
  if (!games) return;
  cudaSetDevice(games->ctx->device);
  cudaStreamSynchronize(games->ctx->stream);
  GamesState& gs = games->gs;
  cudaFree(gs.rec), cudaFree(gs.keys), cudaFree(gs.plies), cudaFree(gs.perspective), cudaFree(gs.threefold), cudaFree(gs.overflow);
  cudaFree(games->d_actions), cudaFree(games->d_results), cudaFree(games->d_applied), cudaFree(games->d_seeds);
  delete games;
}

int pabi_cuda_games_actions(pabi_cuda_games* games, uint32_t* offsets, pabi_move_t* moves, size_t moves_cap) {
  if (!games || !offsets || (moves_cap && !moves)) return -1;
  pabi_cuda_ctx* ctx = games->ctx;
  try {
    CK(cudaSetDevice(ctx->device)); /* before any allocation: the scratch must live on the context's device */
    const size_t n = games->gs.n;
    u32* d_offsets = (u32*)ensure(ctx, S_SMALL_A, (n + 1) * sizeof(u32));
    u16* d_moves = moves_cap ? (u16*)ensure(ctx, S_SMALL_B, moves_cap * sizeof(u16)) : nullptr;
    uint64_t total = 0;
    const int rc = pabi_cuda_generate_moves_device(ctx, games->gs.rec, n, n, d_offsets, d_moves, moves_cap, &total, nullptr);
    if (rc != 0 && rc != -2) return rc;
    CK(cudaMemcpyAsync(offsets, d_offsets, (n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    if (total <= moves_cap && total) CK(cudaMemcpyAsync(moves, d_moves, total * sizeof(u16), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return total > moves_cap ? -2 : 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_games_apply(pabi_cuda_games* games, const pabi_move_t* actions) {
  if (!games || !actions) return -1;
  pabi_cuda_ctx* ctx = games->ctx;
  try {
    CK(cudaSetDevice(ctx->device));
    const size_t n = games->gs.n;
    CK(cudaMemcpyAsync(games->d_actions, actions, n * sizeof(u16), cudaMemcpyHostToDevice, ctx->stream));
    k_games_apply<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, games->gs, games->d_actions);
    CK(cudaGetLastError());
    return games_overflowed(games);
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_games_result(pabi_cuda_games* games, uint8_t* results) {
  if (!games || !results) return -1;
  pabi_cuda_ctx* ctx = games->ctx;
  try {
    CK(cudaSetDevice(ctx->device));
    const size_t n = games->gs.n;
    k_games_result<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, games->gs, games->d_results);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(results, games->d_results, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_games_positions(pabi_cuda_games* games, pabi_position_t* out) {
  if (!games || !out) return -1;
  pabi_cuda_ctx* ctx = games->ctx;
  try {
    CK(cudaSetDevice(ctx->device));
    const size_t n = games->gs.n;
    PositionImage* d_out = (PositionImage*)ensure(ctx, S_AOS_OUT, n * sizeof(PositionImage));
    k_games_unpack<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(games->gs, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, n * sizeof(PositionImage), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_games_play_random(pabi_cuda_games* games, const uint64_t* seeds, uint32_t max_plies, uint8_t* results,
                                uint32_t* plies, pabi_timing_t* timing) {
  if (!games || !seeds || !results || !plies) return -1;
  pabi_cuda_ctx* ctx = games->ctx;
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    const size_t n = games->gs.n;
    CK(cudaMemcpyAsync(games->d_seeds, seeds, n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    timer.begin_device();
    k_games_play_random<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(
        ctx->dt, games->gs, games->d_seeds, max_plies, games->d_results, games->d_applied);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(results, games->d_results, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(plies, games->d_applied, n * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 0);
    return games_overflowed(games);
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_parse_batch(pabi_cuda_ctx* ctx, const char* text, const uint64_t* line_offsets, size_t n, pabi_position_t* out,
                          int8_t* status, pabi_timing_t* timing) {
  if (!ctx || (n && (!text || !line_offsets || !out || !status))) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    const size_t bytes = (size_t)(line_offsets[n] - line_offsets[0]);
    char* d_text = (char*)ensure(ctx, S_STAGING, bytes + 16);
    u64* d_offsets = (u64*)ensure(ctx, S_PREFIX, (n + 1) * sizeof(u64));
    signed char* d_status = (signed char*)ensure(ctx, S_SMALL_A, n);
    PositionImage* d_out = (PositionImage*)ensure(ctx, S_AOS_OUT, n * sizeof(PositionImage));
    std::vector<u64> rel(n + 1);
    for (size_t i = 0; i <= n; i++) rel[i] = line_offsets[i] - line_offsets[0];
    CK(cudaMemcpyAsync(d_text, text + line_offsets[0], bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_offsets, rel.data(), (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    timer.begin_device();
    k_parse_fens<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(ctx->dt, d_text, d_offsets, n, d_out,
                                                                                            d_status);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(out, d_out, n * sizeof(PositionImage), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, bytes + 112 * n);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  } catch (const std::bad_alloc&) {
    ctx->error = "out of host memory";
    return -3;
  }
}

int pabi_cuda_make_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, const pabi_move_t* moves, size_t n,
                               pabi_position_t* out, pabi_timing_t* timing) {
  if (!ctx || (n && (!positions || !moves || !out))) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    upload_roots(ctx, positions, n);
    Level& l0 = ctx->levels[0];
    u16* d_moves = (u16*)ensure(ctx, S_SMALL_A, n * sizeof(u16));
    u64* d_rec = (u64*)ensure(ctx, S_SMALL_B, n * (POS_WORDS + 1) * sizeof(u64)); /* records, then their hashes */
    u64* d_hash = d_rec + n * POS_WORDS;
    PositionImage* d_out = (PositionImage*)ensure(ctx, S_AOS_OUT, n * sizeof(PositionImage));
    CK(cudaMemcpyAsync(d_moves, moves, n * sizeof(u16), cudaMemcpyHostToDevice, ctx->stream));
    timer.begin_device();
    k_make_moves<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(
        ctx->dt, l0.rec, l0.cap, n, d_moves, (const PositionImage*)ctx->scratch[S_STAGING].ptr, d_rec, n, d_hash);
    LAUNCHED();
    k_unpack<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d_rec, n, n, d_hash, d_out);
    LAUNCHED();
    timer.end_device();
    CK(cudaMemcpyAsync(out, d_out, n * sizeof(PositionImage), cudaMemcpyDeviceToHost, ctx->stream));
    timer.finish(timing, 130 * n);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_generate_moves_device(pabi_cuda_ctx* ctx, const uint64_t* device_records, size_t stride, size_t n,
                                    uint32_t* device_offsets, pabi_move_t* device_moves, size_t moves_cap,
                                    uint64_t* total_moves, pabi_timing_t* timing) {
  if (!ctx || !device_records || !device_offsets || !total_moves) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (ctx->list_two_pass) { /* the round-1 path: sizing pass, scan, host round trip for the total, emitting pass */
      u32* counts = (u32*)ensure(ctx, S_COUNTS, (n + 1) * sizeof(u32));
      u64* prefix = (u64*)ensure(ctx, S_PREFIX, (n + 2) * sizeof(u64));
      timer.begin_device();
      if (n) {
        count_and_scan(ctx, device_records, stride, 0, n, counts, prefix);
        CK(cudaMemcpyAsync(ctx->h_result, prefix + n, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (device_moves && ctx->h_result[0] <= moves_cap) {
          launch_moves(ctx, device_records, stride, n, prefix, device_moves, device_offsets);
        } else {
          k_prefix_to_offsets<<<grid_for(ctx, n + 1, 256, 8), 256, 0, ctx->stream>>>(prefix, n, device_offsets);
          LAUNCHED();
        }
      } else {
        CK(cudaMemsetAsync(device_offsets, 0, sizeof(u32), ctx->stream));
        ctx->h_result[0] = 0;
      }
      timer.end_device();
    } else { /* one kernel, one launch: the records are on the device already */
      const ListInput in{nullptr, device_records, stride, ctx->d_bad_roots};
      const size_t tiles = list_tiles(n);
      timer.begin_device();
      if (n) {
        prepare_list_scan(ctx, tiles);
        launch_list_chunk(ctx, in, n, 0, tiles, device_moves, device_moves ? moves_cap : 0, device_offsets);
        CK(cudaMemcpyAsync(ctx->h_result, ctx->d_list_status + tiles - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      } else {
        CK(cudaMemsetAsync(device_offsets, 0, sizeof(u32), ctx->stream));
        ctx->h_result[0] = 0;
      }
      timer.end_device();
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->h_result[0] &= (1ULL << 62) - 1;
    }
    timer.finish(timing, 0);
    *total_moves = ctx->h_result[0];
    if (timing) timing->algorithmic_bytes = 64 * n + 4 * (n + 1) + 2 * *total_moves;
    return *total_moves > moves_cap && device_moves ? -2 : 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_pack_records(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint64_t* device_records,
                           size_t stride) {
  if (!ctx || (n && (!positions || !device_records)) || stride < n) return bad_arg(ctx, "bad argument");
  try {
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    void* staging = ensure(ctx, S_STAGING, n * sizeof(pabi_position_t));
    CK(cudaMemcpyAsync(staging, positions, n * sizeof(pabi_position_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bad_roots, 0, sizeof(u32), ctx->stream));
    k_pack<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>((const PositionImage*)staging, n, device_records, stride, 0, nullptr,
                                                            ctx->d_bad_roots);
    CK(cudaGetLastError());
    u32 bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->d_bad_roots, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) return bad_arg(ctx, "an input record is not a position (each side needs exactly one king, no pawns on the back ranks)");
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_generate_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint32_t* offsets,
                                   pabi_move_t* moves, size_t moves_cap, pabi_timing_t* timing) {
  if (!ctx || !offsets || (n && !positions) || (moves_cap && !moves)) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      offsets[0] = 0;
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    if (n >= 0xFFFFFFFFull / 256) return bad_arg(ctx, "batch too large for 32-bit CSR offsets");
    if (ctx->list_two_pass) {
      upload_roots(ctx, positions, n);
      Level& l0 = ctx->levels[0];
      u32* d_offsets = (u32*)ensure(ctx, S_SMALL_A, (n + 1) * sizeof(u32));
      u16* d_moves = moves_cap ? (u16*)ensure(ctx, S_SMALL_B, moves_cap * sizeof(u16)) : nullptr;
      alloc_level(ctx, l0, n); /* counts/prefix of level 0 are the sizing arrays */
      timer.begin_device();
      count_and_scan(ctx, l0.rec, l0.cap, 0, n, l0.counts, l0.prefix);
      CK(cudaMemcpyAsync(ctx->h_result, l0.prefix + n, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      const size_t total = (size_t)ctx->h_result[0];
      if (total <= moves_cap && total) {
        launch_moves(ctx, l0.rec, l0.cap, n, l0.prefix, d_moves, d_offsets);
      } else {
        k_prefix_to_offsets<<<grid_for(ctx, n + 1, 256, 8), 256, 0, ctx->stream>>>(l0.prefix, n, d_offsets);
        LAUNCHED();
      }
      timer.end_device();
      CK(cudaMemcpyAsync(offsets, d_offsets, (n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
      if (total <= moves_cap && total)
        CK(cudaMemcpyAsync(moves, d_moves, total * sizeof(u16), cudaMemcpyDeviceToHost, ctx->stream));
      timer.finish(timing, 64 * n + 4 * (n + 1) + 2 * total);
      if (total > moves_cap) {
        ctx->error = "moves_cap too small; offsets[n] holds the required size";
        return -2;
      }
      return 0;
    }
    /* The batch in chunks of LIST_CHUNK_TILES tiles, three streams: chunk c's positions travel to the device while the
     * kernel works on chunk c-1 and the lists of chunk c-2 travel back.  The kernel reads the caller's 112-byte records
     * directly and sizes its output itself (decoupled look-back across the chunks), so the host waits for nothing but the
     * per-chunk totals it needs to address the copies back. */
    const size_t tiles = list_tiles(n), chunks = (tiles + LIST_CHUNK_TILES - 1) / LIST_CHUNK_TILES;
    PositionImage* staging = (PositionImage*)ensure(ctx, S_STAGING, n * sizeof(pabi_position_t));
    u32* d_offsets = (u32*)ensure(ctx, S_SMALL_A, (n + 1) * sizeof(u32));
    u16* d_moves = moves_cap ? (u16*)ensure(ctx, S_SMALL_B, moves_cap * sizeof(u16)) : nullptr;
    prepare_list_scan(ctx, tiles);
    while (ctx->list_events.size() < 4 * chunks + 1) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      ctx->list_events.push_back(e);
    }
    if (ctx->list_totals_cap < chunks) {
      if (ctx->h_list_totals) CK(cudaFreeHost(ctx->h_list_totals));
      ctx->h_list_totals = nullptr;
      CK(cudaMallocHost(&ctx->h_list_totals, (chunks + 64) * sizeof(u64)));
      ctx->list_totals_cap = chunks + 64;
    }
    const ListInput in{staging, nullptr, 0, ctx->d_bad_roots};
    cudaEvent_t ready = ctx->list_events[4 * chunks]; /* the scan words are zeroed, the scratch of an earlier call is free */
    CK(cudaMemsetAsync(ctx->d_bad_roots, 0, sizeof(u32), ctx->stream));
    CK(cudaEventRecord(ready, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_in, ready, 0));
    timer.begin_device();
    for (size_t c = 0; c < chunks; c++) {
      const size_t t0 = c * LIST_CHUNK_TILES, t1 = std::min(tiles, t0 + LIST_CHUNK_TILES);
      const size_t p0 = t0 * LIST_TILE, p1 = std::min(n, t1 * LIST_TILE);
      cudaEvent_t arrived = ctx->list_events[4 * c], started = ctx->list_events[4 * c + 1], finished = ctx->list_events[4 * c + 2];
      CK(cudaMemcpyAsync(staging + p0, positions + p0, (p1 - p0) * sizeof(pabi_position_t), cudaMemcpyHostToDevice, ctx->copy_in));
      CK(cudaEventRecord(arrived, ctx->copy_in));
      CK(cudaStreamWaitEvent(ctx->stream, arrived, 0));
      CK(cudaEventRecord(started, ctx->stream));
      launch_list_chunk(ctx, in, n, t0, t1, d_moves, moves_cap, d_offsets);
      CK(cudaEventRecord(finished, ctx->stream));
      CK(cudaMemcpyAsync(ctx->h_list_totals + c, ctx->d_list_status + t1 - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaEventRecord(ctx->list_events[4 * c + 3], ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->h_bad_roots, ctx->d_bad_roots, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream)); /* read in CallTimer::finish */
    timer.end_device();
    size_t done_moves = 0;
    bool fits = true;
    for (size_t c = 0; c < chunks; c++) { /* the lists of a chunk go back as soon as its total is known */
      const size_t t0 = c * LIST_CHUNK_TILES, t1 = std::min(tiles, t0 + LIST_CHUNK_TILES);
      const size_t p0 = t0 * LIST_TILE, p1 = std::min(n, t1 * LIST_TILE);
      CK(cudaEventSynchronize(ctx->list_events[4 * c + 3]));
      const size_t upto = (size_t)(ctx->h_list_totals[c] & ((1ULL << 62) - 1));
      CK(cudaMemcpyAsync(offsets + p0, d_offsets + p0, (p1 - p0 + (p1 == n ? 1 : 0)) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->copy_out));
      fits = fits && upto <= moves_cap;
      if (fits && upto > done_moves)
        CK(cudaMemcpyAsync(moves + done_moves, d_moves + done_moves, (upto - done_moves) * sizeof(u16), cudaMemcpyDeviceToHost, ctx->copy_out));
      done_moves = upto;
    }
    CK(cudaStreamSynchronize(ctx->copy_out));
    timer.finish(timing, 112 * n + 4 * (n + 1) + 2 * done_moves);
    if (timing) { /* device time = the kernels alone: the stream also waited for the copies in */
      timing->device_ms = 0;
      for (size_t c = 0; c < chunks; c++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->list_events[4 * c + 1], ctx->list_events[4 * c + 2]));
        timing->device_ms += ms;
      }
    }
    if (!fits) {
      ctx->error = "moves_cap too small; offsets[n] holds the required size";
      return -2;
    }
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_expand_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n, uint32_t* offsets,
                           pabi_position_t* children, size_t children_cap, pabi_timing_t* timing) {
  if (!ctx || !offsets || (n && !positions)) return bad_arg(ctx, "null argument");
  try {
    CK(cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    if (n == 0) {
      offsets[0] = 0;
      timer.begin_device(), timer.end_device(), timer.finish(timing, 0);
      return 0;
    }
    if (n >= 0xFFFFFFFFull / 256) return bad_arg(ctx, "batch too large for 32-bit CSR offsets");
    upload_roots(ctx, positions, n);
    Level& l0 = ctx->levels[0];
    u32* d_offsets = (u32*)ensure(ctx, S_SMALL_A, (n + 1) * sizeof(u32));
    timer.begin_device();
    count_and_scan(ctx, l0.rec, l0.cap, 0, n, l0.counts, l0.prefix);
    k_prefix_to_offsets<<<grid_for(ctx, n + 1, 256, 8), 256, 0, ctx->stream>>>(l0.prefix, n, d_offsets);
    LAUNCHED();
    CK(cudaMemcpyAsync(offsets, d_offsets, (n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t total = offsets[n];
    if (children && total > children_cap) {
      timer.end_device();
      timer.finish(timing, 0);
      ctx->error = "children_cap too small; offsets[n] holds the required size";
      return -2;
    }
    if (children && total) {
      if (total > ctx->capacity) return bad_arg(ctx, "children exceed the frontier capacity");
      Level& l1 = ctx->levels[1];
      alloc_level(ctx, l1, ctx->capacity);
      launch_expand(ctx, PER_ROOT, frontier_of(l0), 0, n, l0.prefix, frontier_of(l1));
      PositionImage* d_out = (PositionImage*)ensure(ctx, S_AOS_OUT, total * sizeof(PositionImage));
      u64* d_hash = (u64*)ensure(ctx, S_PREFIX, total * sizeof(u64)); /* Position::hash of every child */
      k_child_hashes<<<grid_for(ctx, n, FLAT_THREADS, 4), FLAT_THREADS, FLAT_SMEM, ctx->stream>>>(
          ctx->dt, l0.rec, l0.cap, n, l0.prefix, (const PositionImage*)ctx->scratch[S_STAGING].ptr, d_hash);
      LAUNCHED();
      k_unpack<<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(l1.rec, l1.cap, total, d_hash, d_out);
      LAUNCHED();
      timer.end_device();
      CK(cudaMemcpyAsync(children, d_out, total * sizeof(PositionImage), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      timer.end_device();
    }
    timer.finish(timing, 64 * n + 64 * total);
    return 0;
  } catch (const CudaFail& f) {
    return fail(ctx, f);
  }
}

int pabi_cuda_perft_divide(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, pabi_move_t moves_out[PABI_MAX_MOVES],
                           uint64_t nodes_out[PABI_MAX_MOVES], uint32_t* n_moves, pabi_timing_t* timing) {
  if (!ctx || !root || !moves_out || !nodes_out || !n_moves) return bad_arg(ctx, "null argument");
  if (depth == 0) return bad_arg(ctx, "divide needs depth >= 1");
  uint32_t offsets[2];
  int rc = pabi_cuda_generate_moves_batch(ctx, root, 1, offsets, moves_out, PABI_MAX_MOVES, nullptr);
  if (rc) return rc;
  *n_moves = offsets[1];
  std::vector<pabi_position_t> children(*n_moves ? *n_moves : 1);
  rc = pabi_cuda_expand_batch(ctx, root, 1, offsets, children.data(), children.size(), nullptr);
  if (rc) return rc;
  if (*n_moves == 0) {
    if (timing) std::memset(timing, 0, sizeof *timing);
    return 0;
  }
  return pabi_cuda_perft_batch(ctx, children.data(), *n_moves, (uint8_t)(depth - 1), nodes_out, timing);
}

} /* extern "C" */

/*
 * pabi_cuda.h -- C ABI of libpabi_b200.so, the B200-native replacement for the
 * perft / legal-move-generation path of kirillbobyrev/pabi.
 *
 * pabi has no FFI of its own (pure Rust); the boundary it would bind is its
 * public `pabi::chess` API.  Each entry point below cites the Rust item it
 * replaces (paths relative to the pabi crate root); INTEGRATION.md shows the
 * `extern "C"` block and build.rs hook a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; nothing is retained by the callee except
 *    inside the opaque context;
 *  - status codes: 0 ok; < 0 a parse / validation / argument error (message in
 *    the caller's buffer or pabi_cuda_last_error); > 0 a CUDA error code
 *    (cudaError_t).  There is NO CPU fallback: every move-generation, make-move
 *    and perft entry point runs on the device or fails;
 *  - one context owns one device; calls on one context must be serialised by
 *    the caller, different contexts may be used concurrently;
 *  - positions handed to the device entry points must be positions: built by
 *    pabi_position_from_fen / _parse / _starting or produced by this library.  A
 *    record without exactly one king per side or with a pawn on a back rank
 *    (what the reference's validate() always rejects) makes the call fail with
 *    status -5; other inconsistencies are the caller's responsibility, as they
 *    are for the reference's make_move (src/engine/mod.rs:57-64).
 */
#ifndef PABI_CUDA_H
#define PABI_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Image of `Position` (src/chess/position.rs:45-63) with `Pieces`
 * (src/chess/bitboard.rs:350-357) inlined in field order; 112 bytes. */
typedef struct {
  uint64_t white[6];          /* king, queens, rooks, bishops, knights, pawns */
  uint64_t black[6];
  uint64_t hash;              /* Position::hash (src/chess/position.rs:110): set by the parsers (compute_hash,
                                 :776-806), maintained by every make-move path exactly as make_move does
                                 (:442-466,491-497,...); never read by move generation or perft.  The
                                 reference's piece / en passant keys are build-random (generated.rs:15-16,
                                 build.rs:26-39); here they are fixed and documented (pabi_zobrist_keys) */
  uint16_t fullmove_counter;
  uint8_t castling;           /* src/chess/core.rs:604-612: 1 K, 2 Q, 4 k, 8 q */
  uint8_t side_to_move;       /* src/environment.rs:13-16: 0 white, 1 black */
  uint8_t halfmove_clock;
  uint8_t en_passant_square;  /* 0..63 (A1 = 0, H8 = 63, src/chess/core.rs:169-182), 64 = None */
  uint8_t _pad[2];
} pabi_position_t;

/* `Move` (src/chess/core.rs:24-66): bits 0-5 from, 6-11 to, 12-14 promotion
 * (0 none, 1 knight, 2 bishop, 3 rook, 4 queen; core.rs:700-705). */
typedef uint16_t pabi_move_t;
#define PABI_MAX_MOVES 256 /* MoveList = ArrayVec<Move, 256>, src/chess/core.rs:139-145 */

typedef struct pabi_cuda_ctx pabi_cuda_ctx; /* device, stream, frontier buffers, staged tables */

typedef struct {
  double device_ms;           /* CUDA-event time on the context's stream, first launch to last */
  double host_ms;             /* wall time of the call, copies included */
  uint64_t algorithmic_bytes; /* 2 * 64 B * interior nodes (perft) or 64N + 4(N+1) + 2M (lists) */
  uint64_t kernel_launches;   /* kernels launched by the call */
  uint64_t interior_nodes;    /* positions for which moves were generated */
  /* the fused last-two-plies kernel (K2), which dominates perft: */
  double leaf_kernel_ms;      /* sum of its launches' CUDA-event durations */
  uint64_t leaf_kernel_launches;
  uint64_t leaf_parents;      /* records it read from HBM (ply depth-2) */
  uint64_t leaf_children;     /* positions it generated, counted and never wrote out (ply depth-1) */
  uint64_t merged_records;    /* pabi_cuda_perft_hashed: frontier records removed by transposition merging */
} pabi_timing_t;

/* ---- host-side text API (no device work) -------------------------------- */
/* Position::starting(), src/chess/position.rs:78-91 */
void pabi_position_starting(pabi_position_t* out);
/* Position::from_fen(&str), src/chess/position.rs:157-275 (strict, then validate :934-1032) */
int pabi_position_from_fen(const char* fen, pabi_position_t* out, char* err, size_t err_cap);
/* impl TryFrom<&str> for Position, src/chess/position.rs:809-821 (trim, optional "fen "/"epd " prefix) */
int pabi_position_parse(const char* text, pabi_position_t* out, char* err, size_t err_cap);
/* impl Display for Position, src/chess/position.rs:823-860; returns the length, < 0 if cap is too small */
int pabi_position_to_fen(const pabi_position_t* p, char* out, size_t cap);
/* Move::from_uci / impl Display for Move, src/chess/core.rs:105-134 */
int pabi_move_from_uci(const char* uci, pabi_move_t* out);
int pabi_move_to_uci(pabi_move_t mv, char out[6]);
/* Move::flip_perspective, src/chess/core.rs:91-97 (both squares mirrored over the middle of the board, core.rs:231) */
pabi_move_t pabi_move_flip_perspective(pabi_move_t mv);
/* Position::hash(), src/chess/position.rs:110: the cached field */
uint64_t pabi_position_hash(const pabi_position_t* p);
/* compute_hash, src/chess/position.rs:776-806: the Zobrist hash recomputed from the fields (what the parsers cache) */
uint64_t pabi_position_compute_hash(const pabi_position_t* p);
/* Position::num_pieces(), src/chess/position.rs:122-124 */
uint32_t pabi_position_num_pieces(const pabi_position_t* p);
/* Position::halfmove_clock_expired(), src/chess/position.rs:750-752 */
int pabi_position_halfmove_clock_expired(const pabi_position_t* p);
/* The Zobrist keys (src/chess/generated.rs:7-27): out[0..768) piece keys at player * 384 + kind * 64 + square
 * (get_piece_key), out[768..776) en passant files, out[776] BLACK_TO_MOVE, out[777..781) WHITE_CAN_CASTLE_SHORT,
 * WHITE_CAN_CASTLE_LONG, BLACK_CAN_CASTLE_SHORT, BLACK_CAN_CASTLE_LONG.  The last five are the reference's constants;
 * the first 776 are consecutive splitmix64 outputs from the seed 0x70616269 ("pabi"). */
#define PABI_ZOBRIST_KEYS 781
void pabi_zobrist_keys(uint64_t out[PABI_ZOBRIST_KEYS]);

/* ---- device path --------------------------------------------------------- */
/* device < 0 selects the current CUDA device.  frontier_capacity = positions per
 * breadth-first level buffer (0 = default). */
int pabi_cuda_create(int device, size_t frontier_capacity, pabi_cuda_ctx** out);
void pabi_cuda_destroy(pabi_cuda_ctx* ctx);
const char* pabi_cuda_last_error(pabi_cuda_ctx* ctx);
int pabi_cuda_device(const pabi_cuda_ctx* ctx);

/* Checking builds only (`make -C pabi_b200/csrc bounds`, -DPABI_BOUNDS): the source line of the first out-of-range
 * record or pool index any kernel has computed since the last call (0 = none; the offending access was skipped), or -1
 * when the library was built without the checks (the shipped build). */
int pabi_cuda_bounds_violation(pabi_cuda_ctx* ctx);

/* Page-locked host memory for the arrays of the batched calls below (new; the reference has no counterpart).  Buffers
 * from here are copied to and from the device at the speed of the link instead of through the driver's staging
 * copy of pageable memory (1 M positions in, 27.6 M moves out: 24.7 ms -> 3.4 ms end to end on a B200 box).  Any host pointer is
 * accepted by every call; this is an optimisation only. */
int pabi_cuda_host_alloc(size_t bytes, void** out);
void pabi_cuda_host_free(void* ptr);

/* perft(&Position, depth), src/chess/position.rs:909-924 */
int pabi_cuda_perft(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, uint64_t* nodes,
                    pabi_timing_t* timing);
/* perft with transposition merging ("hashed perft"): after every breadth-first ply the
 * records that are the same position (placement, side to move, castling rights, en passant square) are
 * merged into one record that carries the number of paths reaching it, so shared subtrees are walked
 * once.  Candidates are grouped by a 64-bit tag and then compared in full, so the result is exactly
 * perft(root, depth); only the work changes.  The reference has no such mode (its perft is the plain
 * recursion); throughput of this entry point is "effective" nodes/s and is never what bench.py reports. */
int pabi_cuda_perft_hashed(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth, uint64_t* nodes,
                           pabi_timing_t* timing);
/* The same for n independent roots in one pass (benches/chess.rs:49-104 loops over
 * positions; tests/chess.rs:555-578 loops over a book): nodes_out[i] = perft(roots[i], depth). */
int pabi_cuda_perft_batch(pabi_cuda_ctx* ctx, const pabi_position_t* roots, size_t n, uint8_t depth,
                          uint64_t* nodes_out, pabi_timing_t* timing);
/* The same with a depth per root: nodes_out[i] = perft(roots[i], depths[i]) -- the reference's perft suite
 * (tests/chess.rs:622-818, benches/chess.rs:52-86: every vector has its own depth) as ONE call.  Roots of equal
 * depth are walked together. */
int pabi_cuda_perft_batch_depths(pabi_cuda_ctx* ctx, const pabi_position_t* roots, const uint8_t* depths, size_t n,
                                 uint64_t* nodes_out, pabi_timing_t* timing);
/* Root-subtree sharding for multi-GPU runs: the tree is expanded breadth-first to
 * `split_ply` plies (at most depth - 1), this call then walks only the subtrees of
 * that frontier that belong to shard `shard_index` of `shard_count`.  A subtree's
 * shard is a hash of its root record modulo shard_count, i.e. a property of the
 * position, not of its place in the frontier: every rank can expand the top of the
 * tree on its own, in any order, and the shards are still disjoint and exhaustive.
 * Summing *nodes over all shards gives perft(root, depth). */
int pabi_cuda_perft_shard(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth,
                          uint32_t split_ply, uint32_t shard_index, uint32_t shard_count,
                          uint64_t* nodes, pabi_timing_t* timing);
/* Single-process multi-GPU perft: one context per device (pabi_cuda_create), one host thread per
 * context inside the call, shard i of n_ctx on ctxs[i], host-side u64 sum.  per_ctx_ms (optional,
 * [n_ctx]) receives each device's CUDA-event time.  (bench.py uses one process per GPU and
 * torch.distributed instead; this is what a single Rust process would call.) */
int pabi_cuda_perft_multi(pabi_cuda_ctx* const* ctxs, size_t n_ctx, const pabi_position_t* root, uint8_t depth,
                          uint32_t split_ply, uint64_t* nodes, double* per_ctx_ms);
/* perft "divide": the root's moves in generate_moves order with the node count
 * below each (the per-move breakdown `go perft` prints; engine/mod.rs:176-190). */
int pabi_cuda_perft_divide(pabi_cuda_ctx* ctx, const pabi_position_t* root, uint8_t depth,
                           pabi_move_t moves_out[PABI_MAX_MOVES], uint64_t nodes_out[PABI_MAX_MOVES],
                           uint32_t* n_moves, pabi_timing_t* timing);

/* Position::generate_moves(), src/chess/position.rs:317-408, for n positions.
 * CSR output: the moves of position i are moves[offsets[i] .. offsets[i+1]) in the
 * reference's order.  Returns -2 (and the required size in offsets[n]) when
 * moves_cap is too small. */
int pabi_cuda_generate_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                                   uint32_t* offsets /* [n+1] */, pabi_move_t* moves, size_t moves_cap,
                                   pabi_timing_t* timing);
/* generate_moves().len() for n positions (perft's bulk count, position.rs:914-916) */
int pabi_cuda_count_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                                uint32_t* counts /* [n] */, pabi_timing_t* timing);
/* Position::in_check(), src/chess/position.rs:735-746, for n positions */
int pabi_cuda_in_check_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                             uint8_t* in_check /* [n] */, pabi_timing_t* timing);
/* Position::attack_info(), src/chess/position.rs:285-292 -> AttackInfo (src/chess/attacks.rs:89-97) for n
 * positions: out[5*i .. 5*i+5) = attacks, checkers, pins, xrays, safe_king_squares of position i. */
int pabi_cuda_attack_info_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                                uint64_t* out /* [5 * n] */, pabi_timing_t* timing);
/* Position::make_move(&Move), src/chess/position.rs:414-433: out[i] = positions[i] after moves[i]
 * (moves must be legal).  out[i].hash = positions[i].hash with the XORs make_move applies. */
int pabi_cuda_make_moves_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, const pabi_move_t* moves,
                               size_t n, pabi_position_t* out, pabi_timing_t* timing);
/* One breadth-first ply: every child of every position, in generate_moves order
 * (clone + make_move of position.rs:918-920).  children may be NULL to get only
 * the offsets.  Returns -2 when children_cap is too small.  All children of one call
 * must fit one frontier buffer (pabi_cuda_create's frontier_capacity, default 32 Mi
 * records): a larger batch is an argument error (-1), split it. */
int pabi_cuda_expand_batch(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                           uint32_t* offsets /* [n+1] */, pabi_position_t* children, size_t children_cap,
                           pabi_timing_t* timing);

/* Seeded uniform-random playouts on the device: the position source of the datagen-style workload
 * (the loop a src/datagen driver runs around Position::generate_moves / make_move,
 * src/chess/position.rs:317-433; src/datagen itself is empty upstream).  Game g starts from starts[g];
 * at each ply the move is moves[splitmix64(seeds[g]) % n] in generate_moves order; the position after
 * every stride-th ply (ply 0 included) is written to out[g * per_game_cap + k] until there is no legal
 * move, the halfmove clock reaches 100 (Position::halfmove_clock_expired, position.rs:750-752),
 * max_plies is reached or per_game_cap positions are recorded.  counts[g] = positions recorded. */
int pabi_cuda_random_playouts(pabi_cuda_ctx* ctx, const pabi_position_t* starts, const uint64_t* seeds, size_t n_games,
                              uint32_t max_plies, uint32_t stride, uint32_t per_game_cap, pabi_position_t* out,
                              uint32_t* counts, pabi_timing_t* timing);

/* FEN / EPD book parsing on the device (impl TryFrom<&str> for Position, src/chess/position.rs:809-821, i.e.
 * trim + optional "fen "/"epd " prefix + from_fen :157-275 + validate :934-1032), one thread per line:
 * line i is text[line_offsets[i] .. line_offsets[i+1]) (a trailing '\n' is ignored).  status[i] = 0 and
 * out[i] = the position, or a negative class (-1 placement, -2 side to move, -3 castling, -4 en passant,
 * -5 halfmove clock, -6 fullmove counter, -7 trailing fields, -8 illegal position) and out[i] zeroed;
 * pabi_position_parse on the host gives the reference's message for a rejected line. */
int pabi_cuda_parse_batch(pabi_cuda_ctx* ctx, const char* text, const uint64_t* line_offsets /* [n+1] */, size_t n,
                          pabi_position_t* out /* [n] */, int8_t* status /* [n] */, pabi_timing_t* timing);
/* The 64-byte device record of a position as a wire / on-disk format */
void pabi_position_pack64(const pabi_position_t* p, uint64_t out[8]);
void pabi_position_unpack64(const uint64_t in[8], pabi_position_t* out);

/* ---- Game environment: `Game` (src/chess/game.rs:21-111) for a batch of games ------------------
 * State lives on the device: position, perspective (the root's side to move, game.rs:36), the
 * repetition table (src/chess/zobrist.rs:10-36) and the threefold flag.  The Syzygy adjudication
 * branch of Game::result (game.rs:72-96) needs the third-party shakmaty-syzygy prober and tablebase
 * files and is not provided: results are those of the reference with no tablebase loaded.
 * Result codes: 0 = None, 1 = Win, 2 = Draw, 3 = Loss (GameResult, src/environment.rs:56-61, from
 * the perspective of the player to move at the root). */
typedef struct pabi_cuda_games pabi_cuda_games;
#define PABI_NO_MOVE 0xFFFFu
/* Game::new for n roots (game.rs:30-47); history_cap = most actions a game will ever receive */
int pabi_cuda_games_create(pabi_cuda_ctx* ctx, const pabi_position_t* roots, size_t n, uint32_t history_cap,
                           pabi_cuda_games** out);
void pabi_cuda_games_destroy(pabi_cuda_games* games);
/* Game::actions (game.rs:50-52) of every game, CSR like pabi_cuda_generate_moves_batch */
int pabi_cuda_games_actions(pabi_cuda_games* games, uint32_t* offsets /* [n+1] */, pabi_move_t* moves, size_t moves_cap);
/* Game::apply (game.rs:54-59): actions[g] is applied to game g; PABI_NO_MOVE leaves it untouched */
int pabi_cuda_games_apply(pabi_cuda_games* games, const pabi_move_t* actions /* [n] */);
/* Game::result (game.rs:61-110) */
int pabi_cuda_games_result(pabi_cuda_games* games, uint8_t* results /* [n] */);
/* the current Position of every game */
int pabi_cuda_games_positions(pabi_cuda_games* games, pabi_position_t* out /* [n] */);
/* Uniformly random self-play entirely on the device: in every game, actions()[splitmix64(seed) % len] is
 * applied until result() is Some or max_plies actions were applied.  plies[g] = actions applied by this call. */
int pabi_cuda_games_play_random(pabi_cuda_games* games, const uint64_t* seeds /* [n] */, uint32_t max_plies,
                                uint8_t* results /* [n] */, uint32_t* plies /* [n] */, pabi_timing_t* timing);

/* ---- device-resident variants (inputs/outputs already in HBM) ------------- */
/* Positions as 64-byte device records: the eight words of pabi_position_pack64 in
 * pairs, pair k (words 2k, 2k+1) of record i at records[(k * stride + i) * 2]; `records` holds 8 * stride
 * words and must be 16-byte aligned.  Used by bench.py to time the kernels without copies. */
int pabi_cuda_pack_records(pabi_cuda_ctx* ctx, const pabi_position_t* positions, size_t n,
                           uint64_t* device_records, size_t stride);
int pabi_cuda_generate_moves_device(pabi_cuda_ctx* ctx, const uint64_t* device_records, size_t stride, size_t n,
                                    uint32_t* device_offsets, pabi_move_t* device_moves, size_t moves_cap,
                                    uint64_t* total_moves, pabi_timing_t* timing);

#ifdef __cplusplus
}
#endif
#endif /* PABI_CUDA_H */

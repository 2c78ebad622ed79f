/*
 * abi_smoke.c -- TEST INFRASTRUCTURE.  A plain C11 consumer of include/pabi_cuda.h: asserts the layout
 * of both structs of the boundary at compile time, exercises the host entry points, and, when a CUDA
 * device is present, runs one perft through the header exactly as a foreign-language binding would
 * (rust/ffi.rs mirrors these declarations).  Built and run by tests/test_host_api.py (CPU) and
 * tests/test_gpu_parity.py (GPU).
 */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "pabi_cuda.h"

_Static_assert(sizeof(pabi_position_t) == 112, "Position image");
_Static_assert(offsetof(pabi_position_t, white) == 0, "white");
_Static_assert(offsetof(pabi_position_t, black) == 48, "black");
_Static_assert(offsetof(pabi_position_t, hash) == 96, "hash");
_Static_assert(offsetof(pabi_position_t, fullmove_counter) == 104, "fullmove_counter");
_Static_assert(offsetof(pabi_position_t, castling) == 106, "castling");
_Static_assert(offsetof(pabi_position_t, side_to_move) == 107, "side_to_move");
_Static_assert(offsetof(pabi_position_t, halfmove_clock) == 108, "halfmove_clock");
_Static_assert(offsetof(pabi_position_t, en_passant_square) == 109, "en_passant_square");
_Static_assert(sizeof(pabi_timing_t) == 80, "timing record");
_Static_assert(offsetof(pabi_timing_t, device_ms) == 0 && offsetof(pabi_timing_t, host_ms) == 8, "timing: times");
_Static_assert(offsetof(pabi_timing_t, algorithmic_bytes) == 16 && offsetof(pabi_timing_t, kernel_launches) == 24, "timing: counts");
_Static_assert(offsetof(pabi_timing_t, interior_nodes) == 32 && offsetof(pabi_timing_t, leaf_kernel_ms) == 40, "timing: leaf");
_Static_assert(offsetof(pabi_timing_t, leaf_kernel_launches) == 48 && offsetof(pabi_timing_t, leaf_parents) == 56, "timing: leaf");
_Static_assert(offsetof(pabi_timing_t, leaf_children) == 64 && offsetof(pabi_timing_t, merged_records) == 72, "timing: tail");
_Static_assert(sizeof(pabi_move_t) == 2, "Move is 16 bits");

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      fprintf(stderr, "abi_smoke: %s failed (line %d)\n", #cond, __LINE__); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main(void) {
  puts("layout ok");
  pabi_position_t start, parsed;
  char err[256], fen[128];
  pabi_position_starting(&start);
  CHECK(pabi_position_to_fen(&start, fen, sizeof fen) > 0);
  CHECK(strcmp(fen, "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1") == 0);
  CHECK(pabi_position_from_fen(fen, &parsed, err, sizeof err) == 0);
  CHECK(memcmp(&start, &parsed, sizeof start) == 0);
  CHECK(pabi_position_hash(&start) == pabi_position_compute_hash(&start) && pabi_position_hash(&start) != 0);
  CHECK(pabi_position_num_pieces(&start) == 32 && !pabi_position_halfmove_clock_expired(&start));
  CHECK(pabi_position_from_fen("8/8/8/8/8/8/8/8 w - - 0 1", &parsed, err, sizeof err) != 0 && strstr(err, "illegal position"));
  pabi_move_t mv;
  char uci[6];
  CHECK(pabi_move_from_uci("e7e8q", &mv) == 0 && mv == (52 | (60 << 6) | (4 << 12)));
  CHECK(pabi_move_to_uci(pabi_move_flip_perspective(mv), uci) == 5 && strcmp(uci, "e2e1q") == 0);
  uint64_t keys[PABI_ZOBRIST_KEYS], words[8];
  pabi_zobrist_keys(keys);
  CHECK(keys[776] == 0x9E06BAD39D761293ULL);
  pabi_position_pack64(&start, words);
  pabi_position_unpack64(words, &parsed);
  CHECK(memcmp(&start, &parsed, sizeof start) == 0);
  puts("text api ok");

  pabi_cuda_ctx* ctx = NULL;
  const int rc = pabi_cuda_create(-1, 0, &ctx);
  if (rc != 0) { /* no device: there is no CPU fallback, the device path must refuse */
    CHECK(ctx == NULL);
    puts("no device: device entry points refused");
    return 0;
  }
  uint64_t nodes = 0;
  pabi_timing_t timing;
  CHECK(pabi_cuda_perft(ctx, &start, 5, &nodes, &timing) == 0);
  CHECK(nodes == 4865609ULL && timing.kernel_launches > 0 && timing.device_ms > 0);
  pabi_position_t roots[2];
  const uint8_t depths[2] = {3, 4};
  uint64_t counts[2];
  roots[0] = roots[1] = start;
  CHECK(pabi_cuda_perft_batch_depths(ctx, roots, depths, 2, counts, NULL) == 0 && counts[0] == 8902 && counts[1] == 197281);
  memset(&roots[1], 0, sizeof roots[1]); /* not a position: the call must fail cleanly */
  CHECK(pabi_cuda_perft_batch_depths(ctx, roots, depths, 2, counts, NULL) == -5 && strstr(pabi_cuda_last_error(ctx), "not positions"));
  CHECK(pabi_cuda_perft(ctx, &start, 4, &nodes, NULL) == 0 && nodes == 197281ULL); /* the context is still usable */
  pabi_cuda_destroy(ctx);
  puts("device perft ok");
  return 0;
}

/*
 * drivers.c -- loops around the oracle that the reference does not have: a
 * multi-threaded perft split (CPU baseline on all host cores), batched move
 * lists in CSR form and seeded random playouts (TEST INFRASTRUCTURE; see
 * pabi_oracle.h).  The per-node work is oracle_generate_moves /
 * oracle_make_move, i.e. the reference's algorithm (position.rs:909-924).
 */
#include "oracle_internal.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const oracle_position_t *tasks;
  size_t n_tasks;
  uint8_t depth;
  size_t next;
  uint64_t total;
  pthread_mutex_t lock;
} perft_pool_t;

static void *perft_worker(void *arg) {
  perft_pool_t *pool = (perft_pool_t *)arg;
  uint64_t local = 0;
  for (;;) {
    pthread_mutex_lock(&pool->lock);
    size_t i = pool->next++;
    pthread_mutex_unlock(&pool->lock);
    if (i >= pool->n_tasks) break;
    local += oracle_perft(&pool->tasks[i], pool->depth);
  }
  pthread_mutex_lock(&pool->lock);
  pool->total += local;
  pthread_mutex_unlock(&pool->lock);
  return NULL;
}

/* Same value as oracle_perft(p, depth); the tree is cut at ply 2 and the
 * subtrees are handed to `threads` workers from a shared queue. */
uint64_t oracle_perft_mt(const oracle_position_t *p, uint8_t depth, int threads) {
  if (depth < 4 || threads <= 1) return oracle_perft(p, depth);
  uint16_t m1[256], m2[256];
  uint32_t n1 = oracle_generate_moves(p, m1);
  size_t cap = 256 * 256, n_tasks = 0;
  oracle_position_t *tasks = (oracle_position_t *)malloc(cap * sizeof *tasks);
  if (!tasks) return oracle_perft(p, depth);
  for (uint32_t i = 0; i < n1; i++) {
    oracle_position_t a = *p;
    oracle_make_move(&a, m1[i]);
    uint32_t n2 = oracle_generate_moves(&a, m2);
    for (uint32_t j = 0; j < n2; j++) {
      tasks[n_tasks] = a;
      oracle_make_move(&tasks[n_tasks], m2[j]);
      n_tasks++;
    }
  }
  perft_pool_t pool = {tasks, n_tasks, (uint8_t)(depth - 2), 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  int started = 0;
  for (int t = 0; t < threads; t++)
    if (pthread_create(&tid[started], NULL, perft_worker, &pool) == 0) started++;
  if (started == 0) perft_worker(&pool);
  for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
  free(tasks);
  return pool.total;
}

/* generate_moves (position.rs:317-408) for n positions; CSR output.  Returns the
 * total number of moves (also when it exceeds moves_cap, in which case moves past
 * the cap are not stored). */
uint64_t oracle_generate_moves_batch(const oracle_position_t *p, size_t n, uint32_t *offsets,
                                     uint16_t *moves, size_t moves_cap) {
  uint64_t total = 0;
  uint16_t tmp[256];
  for (size_t i = 0; i < n; i++) {
    offsets[i] = (uint32_t)total;
    uint32_t k = oracle_generate_moves(&p[i], tmp);
    if (total + k <= moves_cap) memcpy(moves + total, tmp, k * sizeof(uint16_t));
    total += k;
  }
  offsets[n] = (uint32_t)total;
  return total;
}

static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* Seeded uniform-random playout: from `start`, each ply pick
 * moves[rng % n] in the reference's move order; records the position after every
 * `stride`-th ply (ply 0 included) until no moves, halfmove clock >= 100
 * (position.rs:750-752), `max_plies` or out_cap.  Returns positions written. */
size_t oracle_random_playout(const oracle_position_t *start, uint64_t seed, int max_plies,
                             oracle_position_t *out, size_t out_cap, int stride) {
  oracle_position_t p = *start;
  uint16_t moves[256];
  size_t written = 0;
  if (stride < 1) stride = 1;
  for (int ply = 0;; ply++) {
    if (ply % stride == 0) {
      if (written >= out_cap) break;
      out[written++] = p;
    }
    if (ply >= max_plies || p.halfmove_clock >= 100) break;
    uint32_t n = oracle_generate_moves(&p, moves);
    if (n == 0) break;
    oracle_make_move(&p, moves[splitmix64(&seed) % n]);
  }
  return written;
}

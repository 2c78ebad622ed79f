/*
 * game.c -- CPU restatement of pabi's `Game` environment (TEST INFRASTRUCTURE; see
 * pabi_oracle.h).  Follows /root/reference/src/chess/game.rs:21-111 (Game::new, actions, apply,
 * result), src/chess/zobrist.rs:10-36 (RepetitionTable::record: true exactly when a key has been
 * seen 3 times) and src/environment.rs:56-74 (GameResult from the perspective of the player to
 * move at the root).  The Syzygy adjudication branch (game.rs:72-96) depends on the third-party
 * shakmaty-syzygy prober and tablebase files and is not restated: a game here is the reference's
 * game with a tablebase of zero pieces.
 *
 * The repetition key is the hash as make_move maintains it (oracle_make_move restates
 * position.rs:414-732 including what it does NOT do: the side-to-move key is never toggled and an
 * en passant key is never removed), so repetitions are detected on placement + castling rights only,
 * exactly as in the reference.
 */
#include "oracle_internal.h"

#include <stdlib.h>
#include <string.h>

struct oracle_game {
  oracle_position_t position;
  int perspective;
  int threefold;
  uint32_t n_keys, cap_keys;
  uint64_t *keys; /* every key recorded so far (a multiset, as the reference's HashMap<Key, u8>) */
  uint16_t moves[256];
  uint32_t n_moves;
};

/* zobrist.rs:28-32 */
static int record(oracle_game_t *g, uint64_t key) {
  int count = 1;
  for (uint32_t i = 0; i < g->n_keys; i++) count += g->keys[i] == key;
  if (g->n_keys == g->cap_keys) {
    g->cap_keys = g->cap_keys ? 2 * g->cap_keys : 64;
    g->keys = (uint64_t *)realloc(g->keys, g->cap_keys * sizeof(uint64_t));
  }
  g->keys[g->n_keys++] = key;
  return count == 3;
}

/* game.rs:30-47 */
oracle_game_t *oracle_game_new(const oracle_position_t *root) {
  oracle_game_t *g = (oracle_game_t *)calloc(1, sizeof *g);
  g->position = *root;
  g->position.hash = oracle_compute_hash(root);
  (void)record(g, g->position.hash);
  g->perspective = root->side_to_move;
  g->n_moves = oracle_generate_moves(&g->position, g->moves);
  return g;
}

void oracle_game_free(oracle_game_t *g) {
  if (g) free(g->keys);
  free(g);
}

/* game.rs:50-52 */
uint32_t oracle_game_actions(const oracle_game_t *g, uint16_t out[256]) {
  memcpy(out, g->moves, g->n_moves * sizeof(uint16_t));
  return g->n_moves;
}

/* game.rs:54-59 */
void oracle_game_apply(oracle_game_t *g, uint16_t action) {
  oracle_make_move(&g->position, action);
  g->threefold = record(g, g->position.hash);
  g->n_moves = oracle_generate_moves(&g->position, g->moves);
}

/* game.rs:61-110 without the tablebase branch.  0 none, 1 win, 2 draw, 3 loss (environment.rs:56-61 order + 1). */
int oracle_game_result(const oracle_game_t *g) {
  if (g->threefold) return 2;
  if (g->position.halfmove_clock >= 100) return 2; /* position.rs:750-752 */
  if (g->n_moves == 0) {
    if (!oracle_in_check(&g->position)) return 2;
    return g->perspective == g->position.side_to_move ? 3 : 1;
  }
  return 0;
}

void oracle_game_position(const oracle_game_t *g, oracle_position_t *out) { *out = g->position; }

static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* Plays uniformly random actions (actions()[splitmix64(seed) % n]) until result() is Some or
 * max_plies actions were applied.  Returns the result code; *plies = actions applied. */
int oracle_game_play_random(oracle_game_t *g, uint64_t seed, uint32_t max_plies, uint32_t *plies) {
  uint32_t ply = 0;
  int r;
  while ((r = oracle_game_result(g)) == 0 && ply < max_plies) {
    oracle_game_apply(g, g->moves[splitmix64(&seed) % g->n_moves]);
    ply++;
  }
  if (plies) *plies = ply;
  return r;
}

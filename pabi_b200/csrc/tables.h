/*
 * tables.h -- host-side construction of the lookup tables the kernels stage
 * into shared memory / read through L1.  Contents follow pabi's generated
 * tables (/root/reference/src/chess/generated.rs:30-82, generated/ directory), rebuilt
 * from their mathematical definition; see tables.cpp.
 */
#pragma once
#include "chess.cuh"

namespace pabi {

struct HostTables {
  SharedTables shared;              /* image copied to every block's shared memory */
  u64 rook_attacks[ROOK_TABLE_SIZE]; /* global memory, read through L1 */
  u64 rays[64 * 64];                /* RAYS, generated/rays.rs */
  u64 zobrist[ZOBRIST_KEYS];        /* Zobrist keys in the layout of chess.cuh (generated.rs:7-27) */
};

/* Built once per process (thread-safe); aborts if a multiply-shift constant
 * fails to index a square's block perfectly. */
const HostTables& host_tables();

/* Reference slider attacks by ray walking (the definition the tables are built
 * from); exposed for the host-side position code (validation needs checkers). */
u64 walk_bishop_attacks(int sq, u64 occ);
u64 walk_rook_attacks(int sq, u64 occ);

/* compute_hash (position.rs:776-806) of a device record, on the host */
u64 host_compute_hash(const Pos& p);

} /* namespace pabi */

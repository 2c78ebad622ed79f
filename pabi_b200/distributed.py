"""Multi-GPU perft: one process per GPU, root-subtree sharding, one u64 reduction.

perft subtrees are independent, so the path has no data exchange while it runs
(SURVEY.md 8e): every rank expands the root breadth-first to `split_ply` on its
own GPU (about a hundred microseconds of redundant work), walks the subtrees whose
root record hashes to its rank (`shard_of`), and the ranks then sum one 64-bit count with
`torch.distributed.all_reduce` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

DEFAULT_SPLIT_PLY = 4  # 197 281 subtrees from the start position: balances 8 GPUs to < 1 %


def world() -> tuple:
    """(rank, world_size, local_rank) from torch.distributed or the torchrun environment."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def all_reduce_sum_u64(value: int, device: Optional[torch.device] = None) -> int:
    """Sum of one unsigned 64-bit count over all ranks.  The count travels as two
    32-bit halves in int64 lanes so that totals up to 2**64 - 1 survive."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    halves = torch.tensor([value & 0xFFFFFFFF, value >> 32], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(halves, op=dist.ReduceOp.SUM)
    lo, hi = (int(x) for x in halves.tolist())
    return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF


def all_reduce_max(values, device: Optional[torch.device] = None):
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def shard_of(record_words, shard_count: int) -> int:
    """The shard a subtree belongs to under `pabi_cuda_perft_shard`: a hash of its root's 64-byte record (the eight words
    of `Position.pack64()`) modulo the shard count -- `shard_hash` of pabi_b200/csrc/kernels.cuh, restated here so that the
    contract is checkable on the host.  A property of the position alone, whatever the order of the frontier."""
    mask = 0xFFFFFFFFFFFFFFFF
    h = 0x9E3779B97F4A7C15
    for w in record_words:
        h = ((h ^ int(w)) * 0xBF58476D1CE4E5B9) & mask
        h ^= h >> 29
    h = (h * 0x94D049BB133111EB) & mask
    return (h ^ (h >> 32)) % shard_count


def shard_ranges(n_items: int, rank: int, world_size: int) -> range:
    """Interleaved assignment used for books / position batches: items rank, rank + world, ..."""
    return range(rank, n_items, world_size)


def perft_distributed(engine, root, depth: int, split_ply: int = DEFAULT_SPLIT_PLY, device: Optional[torch.device] = None) -> int:
    """perft(root, depth) over all ranks of the default process group.  `engine` is a
    :class:`pabi_b200.Engine` bound to this rank's GPU."""
    rank, size, _ = world()
    local = engine.perft(root, depth) if size == 1 else engine.perft_shard(root, depth, split_ply, rank, size)
    return all_reduce_sum_u64(local, device)


def perft_batch_distributed(engine, positions, depth: int, device: Optional[torch.device] = None):
    """Per-position perft counts of a book sharded over the ranks (position i on rank i mod world);
    every rank returns the full array."""
    import numpy as np

    rank, size, _ = world()
    n = len(positions)
    mine = list(shard_ranges(n, rank, size))
    local = engine.perft_batch(positions[mine], depth) if mine else np.zeros(0, dtype=np.uint64)
    out = torch.zeros(2 * n, dtype=torch.int64, device=device or "cpu")
    idx = torch.tensor(mine, dtype=torch.int64, device=out.device)
    vals = torch.from_numpy(local.astype(np.uint64).view(np.int64).copy()).to(out.device)
    out[2 * idx] = vals & 0xFFFFFFFF
    out[2 * idx + 1] = (vals >> 32) & 0xFFFFFFFF
    if size > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)  # disjoint supports: a gather expressed as a sum
    res = out.cpu().numpy().astype(np.uint64)
    return res[0::2] | (res[1::2] << np.uint64(32))

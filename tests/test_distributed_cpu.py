"""The N > 1 host logic on CPU: two gloo ranks shard a perft by root subtree and a book by
position, and reduce the counts (pabi_b200/distributed.py).  The engine is replaced by an
oracle-backed stand-in with the same sharding contract as pabi_cuda_perft_shard, so this checks
the plumbing (rank arithmetic, u64 reduction), not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


class OracleEngine:
    """Same contract as pabi_b200.Engine.perft_shard / perft_batch, computed by the CPU oracle."""

    def perft(self, root, depth):
        return oracle.perft(root, depth)

    def perft_shard(self, root, depth, split_ply, shard_index, shard_count):
        if depth == 0:
            return 1 if shard_index == 0 else 0
        frontier = [root]
        plies = min(split_ply, depth - 1)
        for _ in range(plies):
            nxt = []
            for p in frontier:
                for m in oracle.generate_moves(p):
                    c = oracle.copy(p)
                    oracle.make_move(c, m)
                    nxt.append(c)
            frontier = nxt
        # a subtree's shard is a hash of its root record (pabi_b200.distributed.shard_of), not its place in the frontier
        import pabi_b200 as pb
        from pabi_b200 import distributed as D
        mine = [p for p in frontier if D.shard_of(pb.Position.from_fen(oracle.to_fen(p)).pack64(), shard_count) == shard_index]
        return sum(oracle.perft(p, depth - plies) for p in mine)

    def perft_batch(self, positions, depth):
        return np.array([oracle.perft(oracle.Position.from_buffer_copy(positions[i].tobytes()), depth)
                         for i in range(len(positions))], dtype=np.uint64)


def _worker(rank, world_size, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from pabi_b200 import distributed as D
        engine = OracleEngine()
        kiwipete = oracle.parse("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
        out = {"rank": D.world()[0]}
        out["kiwipete3"] = D.perft_distributed(engine, kiwipete, 3, split_ply=2)
        out["start4"] = D.perft_distributed(engine, oracle.starting(), 4, split_ply=2)
        out["depth1"] = D.perft_distributed(engine, kiwipete, 1, split_ply=4)
        out["big"] = D.all_reduce_sum_u64(0xF000000000000001 if rank == 0 else 0x0FFFFFFFFFFFFFF0)
        book = oracle.random_playout(oracle.starting(), 0xB00C, 40)
        out["book"] = D.perft_batch_distributed(engine, book, 2).tolist()
        out["max"] = D.all_reduce_max([float(rank), 5.0 - rank])
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_shard_and_reduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert sorted(results.keys()) == [0, 1]
    book = oracle.random_playout(oracle.starting(), 0xB00C, 40)
    want_book = [oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), 2) for i in range(len(book))]
    for rank in (0, 1):
        r = results[rank]
        assert r["rank"] == rank
        assert r["kiwipete3"] == 97862 and r["start4"] == 197281 and r["depth1"] == 48
        assert r["big"] == 0xFFFFFFFFFFFFFFF1
        assert r["book"] == want_book
        assert r["max"] == [1.0, 5.0]


def test_single_process_paths_need_no_process_group():
    from pabi_b200 import distributed as D
    assert D.all_reduce_sum_u64(2**63 + 5) == 2**63 + 5
    assert list(D.shard_ranges(10, 1, 4)) == [1, 5, 9]
    assert D.perft_distributed(OracleEngine(), oracle.starting(), 3) == 8902


def test_chess324_stand_in_workload():
    import workloads
    fens = workloads.chess324_fens()
    assert len(fens) == 324 and len(set(fens)) == 324
    assert fens[0] == "rbbnknqr/pppppppp/8/8/8/8/PPPPPPPP/RBBNKNQR w KQkq - 0 1"
    # perft 1..3 of the example array
    p = oracle.parse("rbbnknqr/pppppppp/8/8/8/8/PPPPPPPP/RQNNKBBR w KQkq - 0 1")
    assert [oracle.perft(p, d) for d in (1, 2, 3)] == [20, 400, 9030]
    book = workloads.chess324_synth_book()
    assert 1500 <= len(book) <= 1620

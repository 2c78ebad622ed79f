"""Independent published perft values: oracle (CPU) and CUDA path (GPU)."""
import os

import pytest

import oracle
from published_perft import MAX_MOVES_FEN, PUBLISHED

THREADS = min(8, os.cpu_count() or 1)


@pytest.mark.parametrize("fen,depth,nodes", [c for c in PUBLISHED if c[2] < 50_000_000])
def test_oracle_matches_published_values(fen, depth, nodes):
    assert oracle.perft(oracle.parse(fen), depth, THREADS) == nodes


def test_oracle_max_moves_position():
    assert len(oracle.generate_moves(oracle.parse(MAX_MOVES_FEN))) == 218


@pytest.mark.gpu
def test_gpu_matches_published_values():
    import pabi_b200 as pb
    e = pb.Engine(0)
    for fen, depth, nodes in PUBLISHED:
        assert e.perft(pb.Position.from_fen(fen), depth) == nodes, fen
        # and with a tiny frontier (chunked path)
    small = pb.Engine(0, frontier_capacity=8192)
    for fen, depth, nodes in PUBLISHED:
        if nodes < 5_000_000:
            assert small.perft(pb.Position.from_fen(fen), depth) == nodes, fen
    p = pb.Position.from_fen(MAX_MOVES_FEN)
    moves = p.generate_moves(e)
    assert len(moves) == 218 and [m.raw for m in moves] == oracle.generate_moves(oracle.parse(MAX_MOVES_FEN))
    assert e.perft(p, 3) == oracle.perft(oracle.parse(MAX_MOVES_FEN), 3)
    e.close(), small.close()

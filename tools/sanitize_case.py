"""Small end-to-end exercise of every kernel, for compute-sanitizer (memcheck / racecheck)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402

import pabi_b200 as pb  # noqa: E402

kiwi = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
for cap in (0, 5000):
    e = pb.Engine(0, cap)
    assert e.perft(pb.Position.from_fen(kiwi), 4) == 4085603
    assert e.perft(pb.Position.starting(), 5) == 4865609
    assert sum(e.perft_shard(pb.Position.from_fen(kiwi), 4, 2, i, 3) for i in range(3)) == 4085603
    batch = oracle.random_playout(oracle.starting(), 0x5EED0001, 150)
    got = e.perft_batch(batch, 3)
    want = [oracle.perft(oracle.Position.from_buffer_copy(batch[i].tobytes()), 3) for i in range(len(batch))]
    assert got.tolist() == want
    off, mv = e.generate_moves_batch(batch)
    woff, wmv = oracle.generate_moves_batch(batch)
    assert np.array_equal(off, woff) and np.array_equal(mv, wmv)
    off, ch = e.expand_batch(batch)
    assert np.array_equal(off, woff)
    e.in_check_batch(batch)
    e.count_moves_batch(batch)
    assert sum(n for _, n in e.divide(pb.Position.from_fen(kiwi), 3)) == 97862
    e.close()
print("sanitize case ok")

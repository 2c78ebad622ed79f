"""One-off differential soak (not part of pytest): perft(4) and perft(5) per root for samples of playout positions, GPU (per-root
accounting through the speculative and virtual plies) against the oracle on all host threads."""
import ctypes as C, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle, workloads
import pabi_b200 as pb
e = pb.Engine(0)
small = pb.Engine(0, 50_000)
threads = os.cpu_count() or 1
positions = workloads.playout_positions(400_000, 0x50AC0000)
for depth, step, count in ((4, 37, 8000), (5, 401, 600)):
    sample = np.ascontiguousarray(positions[3::step][:count])
    t0 = time.time()
    got = e.perft_batch(sample, depth)
    got_small = small.perft_batch(sample, depth)
    t1 = time.time()
    want = np.array([oracle.perft(oracle.Position.from_buffer_copy(sample[i].tobytes()), depth, threads) for i in range(len(sample))], dtype=np.uint64)
    assert np.array_equal(got, want) and np.array_equal(got_small, want), np.nonzero(got != want)[0][:5]
    print(f"perft({depth}) x {len(sample)} roots bit-exact ({int(want.sum())} nodes; GPU {t1 - t0:.2f} s incl. the 50 000-record engine, oracle {time.time() - t1:.1f} s)", flush=True)
print("soak ok")

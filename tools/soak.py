"""One-off differential soak on cuda:0 (not part of pytest): large random samples, CUDA path vs oracle.
Usage: python tools/soak.py [positions [seed]]   (default 3 000 000 positions of the documented seed)

  - 3M playout positions (startpos + all Chess324 arrays): ordered move lists, counts, in_check
  - 200k of them: perft(3) per root
  - 20k of them: every child of the position, field for field
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
import workloads  # noqa: E402

import pabi_b200 as pb  # noqa: E402

e = pb.Engine(0)
t0 = time.time()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
SEED = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x5EED0000
positions = workloads.playout_positions(N, SEED)
print(f"{len(positions)} positions generated in {time.time() - t0:.1f}s", flush=True)
off, mv = e.generate_moves_batch(positions)
woff, wmv = oracle.generate_moves_batch(positions)
assert np.array_equal(off, woff) and np.array_equal(mv, wmv)
assert np.array_equal(e.count_moves_batch(positions), np.diff(woff))
print(f"move lists: {len(mv)} moves bit-exact", flush=True)

threads = os.cpu_count() or 1
sample = np.ascontiguousarray(positions[::15][:200_000])
got = e.perft_batch(sample, 3)
lib = oracle.lib()
import ctypes as C
want = np.zeros(len(sample), dtype=np.uint64)
for i in range(len(sample)):
    want[i] = lib.oracle_perft(C.cast(sample.ctypes.data + 112 * i, C.POINTER(oracle.Position)), 3)
assert np.array_equal(got, want), np.nonzero(got != want)[0][:5]
print(f"perft(3) x {len(sample)} roots bit-exact ({int(want.sum())} nodes)", flush=True)

sample = np.ascontiguousarray(positions[7::150][:20_000])
off, children = e.expand_batch(sample)
woff, wmv = oracle.generate_moves_batch(sample)
assert np.array_equal(off, woff)
buf = np.zeros(len(children), dtype=oracle.POSITION_DTYPE)
k = 0
for i in range(len(sample)):
    for m in wmv[woff[i]:woff[i + 1]]:
        C.memmove(buf.ctypes.data + 112 * k, sample.ctypes.data + 112 * i, 112)
        lib.oracle_make_move(C.cast(buf.ctypes.data + 112 * k, C.POINTER(oracle.Position)), int(m))
        k += 1
assert buf.tobytes() == children.tobytes()  # all 112 bytes, Position::hash included
print(f"children: {len(children)} positions equal field for field (hash included)", flush=True)
import brute  # the independent mailbox generator: move sets and in_check of a 600 000-position sample
sample = np.ascontiguousarray(positions[::5][:600_000])
soff, smv = e.generate_moves_batch(sample)
bad, first, total = brute.compare_batch(sample, soff, smv, e.in_check_batch(sample))
assert bad == 0 and total == len(smv), first
print(f"brute-force generator: {total} moves of {len(sample)} positions agree as sets, in_check agrees", flush=True)
ic = e.in_check_batch(positions[:500_000])
ai = e.attack_info_batch(positions[:500_000])
assert np.array_equal(ic, ai[:, 1] != 0)
print("soak ok", flush=True)

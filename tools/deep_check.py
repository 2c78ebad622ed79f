"""Development helper (under gpurun): perft(9) of the start position (2 439 530 234 167) and a large per-root batch, both
through the speculative / virtual plies with chunking."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import pabi_b200 as pb
from pabi_b200 import workloads as W
e = pb.Engine(0)
t0 = time.perf_counter()
n = e.perft(pb.Position.starting(), 9)
print("perft(9)", n, n == 2439530234167, round(e.last_timing.device_ms, 1), "ms", e.last_timing.kernel_launches, "launches", round(n / e.last_timing.device_ms / 1e9, 3), "T nodes/s")
positions = W.playout_positions(e, 200_000)
a = e.perft_batch(positions, 3)
ms = e.last_timing.device_ms
import os
os.environ["PABI_NO_VIRTUAL_PLY"] = "1"; os.environ["PABI_EXACT_LEVELS"] = "1"
e2 = pb.Engine(0)
b = e2.perft_batch(positions, 3)
print("batch 200k x perft(3)", int(a.sum()), np.array_equal(a, b), round(ms, 2), "ms vs", round(e2.last_timing.device_ms, 2), "ms exact path")

"""Development helper (under gpurun): the 1 M move-list batch through the fused K3 path and the round-1 two-pass path."""
import json, sys, time, hashlib
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pabi_b200 as pb
from pabi_b200 import workloads as W
e = pb.Engine(0)
fx = json.loads((ROOT / "tests/golden/workload_fixtures.json").read_text())["playouts"]
sha = lambda a: hashlib.sha256(a.tobytes()).hexdigest()
positions = W.playout_positions(e, 1_000_000)
assert sha(positions) == fx["sha256"]
pin = pb.pinned_empty(len(positions), positions.dtype); pin[:] = positions
po = pb.pinned_empty(len(positions) + 1, np.uint32); pm = pb.pinned_empty(40 * len(positions), np.uint16)
for name, args in (("pageable", (positions,)), ("pinned", (pin, po, pm))):
    best = None
    for _ in range(6):
        t0 = time.perf_counter()
        offsets, moves = e.generate_moves_batch(*args)
        wall = (time.perf_counter() - t0) * 1e3
        ok = sha(offsets) == fx["offsets_sha256"] and sha(moves) == fx["moves_sha256"]
        cur = (round(e.last_timing.device_ms, 4), round(wall, 3), ok, e.last_timing.kernel_launches)
        best = cur if best is None or cur[1] < best[1] else best
    print(name, best, flush=True)
# small and ragged batches
for n in (1, 31, 1023, 1024, 1025, 5000):
    o, m = e.generate_moves_batch(positions[:n])
    o2, m2 = e.generate_moves_batch(positions[:n], None, np.empty(10, dtype=np.uint16)) if n > 1 else (o, m)
    print(n, len(m), int(o[-1]), len(m2) == len(m))

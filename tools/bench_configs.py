"""Times the other BASELINE.json configs on cuda:0 (they are parity-test cases, not bench lines;
this tool records their throughput for DESIGN.md).  Every result is checked against the oracle
or a known answer before it is printed."""
import json
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


def best(fn, reps=5):
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        out.append((e.last_timing.device_ms, (time.perf_counter() - t0) * 1e3, r))
    return min(out, key=lambda x: x[0])


# configs[0]: startpos perft(6)
dev, wall, n = best(lambda: e.perft(pb.Position.starting(), 6))
assert n == 119060324
print(json.dumps({"config": "startpos perft(6)", "nodes": n, "device_ms": round(dev, 4), "e2e_ms": round(wall, 4),
                  "gnps_device": round(n / dev / 1e6, 1), "gnps_e2e": round(n / wall / 1e6, 1)}))

# configs[1]: the reference's perft suite (tests/chess.rs + benches/chess.rs), one root at a time
vec = json.loads((ROOT / "tests/golden/reference_vectors.json").read_text())
cases = [c for c in vec["perft"] + vec["bench_perft"]]
tot_nodes = tot_dev = tot_wall = 0
for c in cases:
    p = pb.Position.starting() if c["fen"] == "startpos" else pb.Position.try_from(c["fen"])
    dev, wall, n = best(lambda: e.perft(p, c["depth"]), reps=3)  # best of 3: the first call of a shape pays lazy module loads
    assert n == c["nodes"], c
    tot_nodes += n
    tot_dev += dev
    tot_wall += wall
print(json.dumps({"config": f"reference perft suite ({len(cases)} vectors incl. Kiwipete d5, CPW d7)", "nodes": tot_nodes,
                  "device_ms": round(tot_dev, 3), "e2e_ms": round(tot_wall, 3), "gnps_device": round(tot_nodes / tot_dev / 1e6, 1)}))

# configs[2]: batched perft(4) over the Chess324 stand-in book
book = workloads.chess324_synth_book()
dev, wall, counts = best(lambda: e.perft_batch(book, 4))
want = [oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), 4, 8) for i in range(len(book))]
assert counts.tolist() == want
n = int(counts.sum())
print(json.dumps({"config": f"perft(4) x {len(book)} chess324-synth positions", "nodes": n, "device_ms": round(dev, 4),
                  "e2e_ms": round(wall, 4), "gnps_device": round(n / dev / 1e6, 1), "gnps_e2e": round(n / wall / 1e6, 1)}))

# configs[4]: ordered move lists for 1M playout positions
positions = workloads.playout_positions(1_000_000)
dev, wall, (offsets, moves) = best(lambda: e.generate_moves_batch(positions), reps=3)
want_offsets, want_moves = oracle.generate_moves_batch(positions)
assert np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves)
t = e.last_timing
print(json.dumps({"config": "move lists, 1M playout positions", "moves": int(len(moves)), "device_ms": round(dev, 4),
                  "e2e_ms": round(wall, 3), "positions_per_s_device": round(1e6 / dev * 1e3), "algorithmic_bytes": t.algorithmic_bytes,
                  "achieved_gbs": round(t.algorithmic_bytes / dev / 1e6, 1)}))
# the same call with every host array in page-locked memory (pabi_cuda_host_alloc)
pinned_in = pb.pinned_empty(len(positions), positions.dtype)
pinned_in[:] = positions
pinned_offsets = pb.pinned_empty(len(positions) + 1, np.uint32)
pinned_moves = pb.pinned_empty(40 * len(positions), np.uint16)
dev, wall, (offsets, moves) = best(lambda: e.generate_moves_batch(pinned_in, pinned_offsets, pinned_moves), reps=3)
assert np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves)
print(json.dumps({"config": "move lists, 1M playout positions, host arrays page-locked", "device_ms": round(dev, 4), "e2e_ms": round(wall, 3),
                  "h2d_bytes": int(pinned_in.nbytes), "d2h_bytes": int(offsets.nbytes + moves.nbytes),
                  "positions_per_s_e2e": round(1e6 / wall * 1e3)}))
t0 = time.perf_counter()
oracle.generate_moves_batch(positions)
print(json.dumps({"config": "move lists, 1M playout positions, oracle on 1 host core", "ms": round((time.perf_counter() - t0) * 1e3, 1)}))

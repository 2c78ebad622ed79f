"""Per-vector device time of the reference perft suite (tests/golden/reference_vectors.json) on cuda:0,
best of --reps runs each; every count is checked.  Used to find positions a kernel change slowed down."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pabi_b200 as pb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--min-nodes", type=int, default=0)
args = ap.parse_args()
e = pb.Engine(0)
vec = json.loads((ROOT / "tests/golden/reference_vectors.json").read_text())
total = 0.0
for c in vec["perft"] + vec["bench_perft"]:
    if c["nodes"] < args.min_nodes:
        continue
    p = pb.Position.starting() if c["fen"] == "startpos" else pb.Position.try_from(c["fen"])
    best = None
    for _ in range(args.reps):
        n = e.perft(p, c["depth"])
        assert n == c["nodes"], c
        t = e.last_timing
        if best is None or t.device_ms < best[0]:
            best = (t.device_ms, t.leaf_kernel_ms, t.kernel_launches)
    total += best[0]
    print(json.dumps({"fen": c["fen"], "depth": c["depth"], "nodes": c["nodes"], "device_ms": round(best[0], 4),
                      "leaf_ms": round(best[1], 4), "launches": best[2], "gnps": round(c["nodes"] / best[0] / 1e6, 1)}))
print(json.dumps({"total_device_ms": round(total, 3)}))

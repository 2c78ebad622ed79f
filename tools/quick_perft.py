"""Development helper: time a few perfts on cuda:0 and print timing records."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pabi_b200 as pb

KNOWN = {5: 4865609, 6: 119060324, 7: 3195901860, 8: 84998978956}
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 0
e = pb.Engine(0, cap)
p = pb.Position.starting()
for depth in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "5,6,7").split(",")]:
    for rep in range(2):
        n = e.perft(p, depth)
        t = e.last_timing
        print(json.dumps({"depth": depth, "nodes": n, "ok": n == KNOWN.get(depth, n), "device_ms": round(t.device_ms, 3),
                          "host_ms": round(t.host_ms, 3), "gnps": round(n / t.device_ms / 1e6, 2),
                          "launches": t.kernel_launches, "interior": t.interior_nodes}), flush=True)

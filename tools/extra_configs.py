"""Times the paths either side of perft on cuda:0 for DESIGN.md: perft with transposition merging,
random self-play of a batch of games, seeded playouts, the device FEN parser.  Every result is checked."""
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
KNOWN = {7: 3195901860, 8: 84998978956, 9: 2439530234167}
start = pb.Position.starting()
for depth in (7, 8, 9):
    best = None
    for _ in range(3):
        n = e.perft_hashed(start, depth)
        assert n == KNOWN[depth]
        t = e.last_timing
        best = t.device_ms if best is None else min(best, t.device_ms)
    print(json.dumps({"config": f"startpos perft({depth}) with transposition merging", "nodes": KNOWN[depth], "device_ms": round(best, 3),
                      "effective_gnps": round(KNOWN[depth] / best / 1e6, 1), "merged_records": e.last_timing.merged_records}), flush=True)

n_games = 65536
roots = np.repeat(workloads.playout_positions(1)[:1], n_games)  # every game from the start position
games = pb.Games(e, roots, history_cap=640)
t0 = time.perf_counter()
results, plies = games.play_random(np.arange(n_games, dtype=np.uint64) + 0xC0FFEE, 600)
wall = (time.perf_counter() - t0) * 1e3
print(json.dumps({"config": f"{n_games} games of uniformly random self-play (Game::apply / result on the device, <= 600 plies)",
                  "plies": int(plies.sum()), "device_ms": round(games.last_timing.device_ms, 2), "e2e_ms": round(wall, 2),
                  "finished": sum(r is not None for r in results)}), flush=True)
games.close()

lines = [oracle.to_fen(oracle.Position.from_buffer_copy(p.tobytes())) for p in workloads.playout_positions(200_000)]
t0 = time.perf_counter()
parsed, status = e.parse_batch(lines)
wall = (time.perf_counter() - t0) * 1e3
assert (status == 0).all()
print(json.dumps({"config": "200 000 FEN lines parsed on the device", "device_ms": round(e.last_timing.device_ms, 3), "e2e_ms": round(wall, 2)}), flush=True)

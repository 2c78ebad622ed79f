"""Development helper: the per-rank device time of startpos perft(8) sharded 8 ways, emulated on ONE GPU by running the
eight shards one after the other, for several split plies (the multi-GPU step time is the max over the shards)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pabi_b200 as pb

e = pb.Engine(0)
p = pb.Position.starting()
shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
e.perft(p, 7)
for split in (2, 3, 4, 5):
    best = None
    for rep in range(3):
        times, total, leaf = [], 0, []
        for i in range(shards):
            total += e.perft_shard(p, 8, split, i, shards)
            times.append(e.last_timing.device_ms)
            leaf.append(e.last_timing.leaf_kernel_ms)
        assert total == 84998978956
        if best is None or max(times) < max(best):
            best, best_leaf = times, leaf
    print(json.dumps({"shards": shards, "split_ply": split, "max_ms": round(max(best), 3), "min_ms": round(min(best), 3),
                      "mean_ms": round(sum(best) / len(best), 3), "leaf_max_ms": round(max(best_leaf), 3),
                      "leaf_mean_ms": round(sum(best_leaf) / len(best_leaf), 3), "launches": e.last_timing.kernel_launches}), flush=True)

"""Development helper (under gpurun, no torch): exactness and perft(8) time of the library currently in place.
Exit code 0 when every count is right and the best perft(8) time is below the threshold in ms (argv[1])."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pabi_b200 as pb

limit = float(sys.argv[1]) if len(sys.argv) > 1 else 1e9
e = pb.Engine(0)
small = pb.Engine(0, 20000)
start = pb.Position.starting()
kiwi = pb.Position.try_from("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
dense = pb.Position.try_from("knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1")
promo = pb.Position.try_from("n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1")
ok = True
out = {}


def check(name, got, want):
    global ok
    out[name] = got if got == want else [got, want]
    ok &= got == want


check("start_d4", e.perft(start, 4), 197281)
check("start_d5", e.perft(start, 5), 4865609)
check("kiwi_d4", e.perft(kiwi, 4), 4085603)
check("start_d6", e.perft(start, 6), 119060324)
check("start_d7", e.perft(start, 7), 3195901860)
check("kiwi_d5", e.perft(kiwi, 5), 193690690)
check("dense_d4", e.perft(dense, 4), 169711264)
check("promo_d6", e.perft(promo, 6), 71179139)
check("small_kiwi_d5", small.perft(kiwi, 5), 193690690)
check("small_dense_d3", small.perft(dense, 3), 1664938)
check("shards_d7", sum(e.perft_shard(start, 7, 4, i, 3) for i in range(3)), 3195901860)
check("batch_d4", e.perft_batch([start, kiwi, dense, promo], 4).tolist(), [197281, 4085603, 169711264, 182838])
check("divide_kiwi_d4", sum(n for _, n in e.divide(kiwi, 4)), 4085603)
if hasattr(e, "perft_hashed"):
    check("hashed_d7", e.perft_hashed(start, 7), 3195901860)
times = []
for _ in range(4):
    check("start_d8", e.perft(start, 8), 84998978956)
    times.append(round(e.last_timing.device_ms, 3))
out["d8_ms"] = times
out["ok"] = ok
print(json.dumps(out), flush=True)
sys.exit(0 if ok and min(times) < limit else 1)

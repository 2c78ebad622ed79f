import json, sys
sys.path.insert(0, ".")
import pabi_b200 as pb
e = pb.Engine(0)
p = pb.Position.starting()
out = {}
for d in (6, 8, 8):
    n = e.perft(p, d); out[f"d{d}"] = (n, round(e.last_timing.device_ms, 3))
s = sum(e.perft_shard(p, 7, 4, i, 3) for i in range(3))
out["shards_d7"] = (s, s == 3195901860)
k = pb.Position.try_from("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
out["kiwipete_d5"] = (e.perft(k, 5), 193690690)
try:
    e.perft_shard(p, 40, 39, 0, 2); out["deep_split"] = "no error"
except pb.PabiCudaError as x:
    out["deep_split"] = str(x)[:120]
out["after_error_d6"] = e.perft(p, 6)
print(json.dumps(out))

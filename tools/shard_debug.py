import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pabi_b200 as pb
e = pb.Engine(0)
p = pb.Position.from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
for depth, split in ((4, 1), (4, 2), (5, 2), (3, 5), (1, 2), (2, 1), (3, 1), (3, 2)):
    whole = e.perft(p, depth)
    for count in (2, 3, 8):
        parts = [e.perft_shard(p, depth, split, i, count) for i in range(count)]
        print(depth, split, count, whole, sum(parts), "OK" if sum(parts) == whole else "BAD", parts if count < 4 else "")

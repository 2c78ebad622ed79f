import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pabi_b200 as pb
e = pb.Engine(0)
for fen, d in (("startpos", 7), ("rnb1kbnr/pp1pp1pp/1qp2p2/8/Q1P5/N7/PP1PPPPP/1RB1KBNR b Kkq - 2 4", 6), ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 5),
               ("r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1", 6)):
    p = pb.Position.starting() if fen == "startpos" else pb.Position.try_from(fen)
    n = e.perft(p, d)
    t = e.last_timing
    print(fen[:30], d, n, round(t.device_ms, 3), "ms", round(n / t.device_ms / 1e9, 3), "T/s", "parents", t.leaf_parents, "children", t.leaf_children, flush=True)

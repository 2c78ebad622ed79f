"""Command line front-end of the GPU perft path.

    python -m pabi_b200 perft --depth 6 [--fen FEN] [--divide]
    python -m pabi_b200 bench [--depth 7]

`bench` prints the OpenBench line pabi's `openbench()` is meant to print ("<nodes> nodes <nps> nps";
src/engine/mod.rs:176-190, tests/integration.rs:23-34 upstream, where it is still `todo!()`), computed
by perft over the positions of the reference's perft benchmark (benches/chess.rs:52-86).
"""
import argparse
import sys
import time

from . import Engine, Position
from .chess import STARTING_FEN

BENCH_POSITIONS = [  # benches/chess.rs:52-86 (position, depth)
    (STARTING_FEN, 5), (STARTING_FEN, 6),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 6),
    ("r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1", 6),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", 5),
    ("r1bqkbnr/pppppppp/2n5/8/3P4/8/PPP1PPPP/RNBQKBNR w KQkq - 1 2", 6),
    ("rnbqkbnr/pppppppp/8/8/8/N7/PPPPPPPP/R1BQKBNR b KQkq - 1 1", 6),
]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="pabi_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    pf = sub.add_parser("perft", help="perft(depth) of a position on the GPU")
    pf.add_argument("--fen", default=STARTING_FEN)
    pf.add_argument("--depth", type=int, required=True)
    pf.add_argument("--divide", action="store_true", help="per-move breakdown")
    pf.add_argument("--device", type=int, default=-1)
    bn = sub.add_parser("bench", help="OpenBench-style '<nodes> nodes <nps> nps' line")
    bn.add_argument("--device", type=int, default=-1)
    args = ap.parse_args(argv)
    engine = Engine(args.device)
    if args.cmd == "perft":
        position = Position.try_from(args.fen)
        if args.divide:
            total = 0
            for move, nodes in engine.divide(position, args.depth):
                print(f"{move}: {nodes}")
                total += nodes
            print(f"\nNodes searched: {total}")
        else:
            nodes = engine.perft(position, args.depth)
            t = engine.last_timing
            print(f"perft({args.depth}) = {nodes}  [{t.device_ms:.3f} ms on device, {nodes / t.device_ms / 1e6:.1f} G nodes/s]")
        return 0
    nodes = 0
    t0 = time.perf_counter()
    for fen, depth in BENCH_POSITIONS:
        nodes += engine.perft(Position.from_fen(fen), depth)
    dt = time.perf_counter() - t0
    print(f"{nodes} nodes {int(nodes / dt)} nps")
    return 0


if __name__ == "__main__":
    sys.exit(main())

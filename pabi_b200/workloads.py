"""Workloads of the perft / move-generation path, produced ON THE DEVICE through the library.

The reference's book (`books/Chess324_xxl_big_+090_+119.epd`) and its 100 000-position FEN file
(`tests/data/positions.fen`) are absent from the mount this repository was built against
(`.MISSING_LARGE_BLOBS`), so SURVEY.md 8(d) defines seeded stand-ins.  Here they are generated with
`pabi_cuda_random_playouts` (bit-exact with the CPU driver of the test-suite, see
tests/test_gpu_parity.py); `load_positions` reads the real files when a caller has them.
"""
from __future__ import annotations

import itertools
from pathlib import Path
from typing import List, Tuple

import numpy as np

from ._native import POSITION_DTYPE
from .chess import Engine, Position

BOOK_SEED = 0xC3240000      # game `index` of the stand-in book plays from seed BOOK_SEED + index
PLAYOUT_SEED = 0x5EED0000   # game g of the move-list workload plays from seed PLAYOUT_SEED + g


def chess324_arrays() -> List[str]:
    """The 18 back-rank arrays of Chess324 for one side: king on e, rooks on a/h, {Q,B,B,N,N} on b,c,d,f,g with
    opposite-coloured bishops, sorted by ASCII ('BBNNQ' first)."""
    out = []
    for perm in sorted(set(itertools.permutations("BBNNQ"))):
        files = dict(zip("bcdfg", perm))
        bishops = [f for f, p in files.items() if p == "B"]
        if (ord(bishops[0]) - ord("a")) % 2 == (ord(bishops[1]) - ord("a")) % 2:
            continue
        out.append("R" + files["b"] + files["c"] + files["d"] + "K" + files["f"] + files["g"] + "R")
    assert len(out) == 18
    return out


def chess324_fens() -> List[str]:
    arrays = chess324_arrays()
    return [f"{b.lower()}/pppppppp/8/8/8/8/PPPPPPPP/{w} w KQkq - 0 1" for w in arrays for b in arrays]


def _positions(fens) -> np.ndarray:
    arr = np.zeros(len(fens), dtype=POSITION_DTYPE)
    for i, fen in enumerate(fens):
        arr[i] = np.frombuffer(bytes(Position.from_fen(fen).raw), dtype=POSITION_DTYPE)[0]
    return arr


def chess324_synth_book(engine: Engine) -> np.ndarray:
    """'chess324-synth' (SURVEY.md 8d, C3 stand-in): every Chess324 start array followed by a seeded 8-ply random
    playout, keeping the positions after plies 0, 2, 4, 6, 8 -> 1 620 positions."""
    starts = _positions(chess324_fens())
    games = engine.random_playouts(starts, [BOOK_SEED + i for i in range(len(starts))], 8, 2)
    return np.ascontiguousarray(np.concatenate(games))


def playout_positions(engine: Engine, target: int, seed: int = PLAYOUT_SEED, batch_games: int = 4096) -> np.ndarray:
    """C5: positions of seeded random playouts from the start position (even games) or a Chess324
    array (odd games, array (g // 2) % 324), every position recorded, games capped at 200 plies, cut at `target`."""
    book = _positions(chess324_fens())
    start = _positions([str(Position.starting())])[0]
    chunks, total, g0 = [], 0, 0
    while total < target:
        ids = np.arange(g0, g0 + batch_games)
        starts = book[(ids // 2) % 324].copy()
        starts[ids % 2 == 0] = start
        for game in engine.random_playouts(starts, (seed + ids).astype(np.uint64), 200, 1):
            chunks.append(game)
            total += len(game)
            if total >= target:
                break
        g0 += batch_games
    return np.ascontiguousarray(np.concatenate(chunks)[:target])


def load_positions(path, engine: Engine) -> Tuple[np.ndarray, np.ndarray]:
    """An EPD book or FEN file (one position per line, parsed like `Position::try_from`, which is how the reference
    reads `books/*.epd` and `tests/data/positions.fen`, tests/chess.rs:556-557) -> (positions, status) with the lines
    parsed on the device (`pabi_cuda_parse_batch`).  EPD operations after the fourth field are not accepted by the
    reference's parser either; such lines come back with a negative status."""
    lines = [line for line in Path(path).read_text().splitlines() if line.strip()]
    return engine.parse_batch(lines)

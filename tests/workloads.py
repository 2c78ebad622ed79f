"""Seeded synthetic workloads of SURVEY.md 8(d), shared by the tests and tools.

The reference's book (books/Chess324_xxl_big_+090_+119.epd) and its 100 000-position
FEN file are absent from the mount, so documented stand-ins are generated with the
oracle: TEST / BENCH INFRASTRUCTURE ONLY."""
import itertools

import numpy as np

import oracle


def chess324_arrays():
    """The 18 back-rank arrays of Chess324 for one side: king on e, rooks on a/h, {Q,B,B,N,N} on
    b,c,d,f,g with opposite-coloured bishops, sorted by ASCII ('BBNNQ' first)."""
    out = []
    for perm in sorted(set(itertools.permutations("BBNNQ"))):
        files = dict(zip("bcdfg", perm))
        bishops = [f for f, p in files.items() if p == "B"]
        if (ord(bishops[0]) - ord("a")) % 2 == (ord(bishops[1]) - ord("a")) % 2:
            continue
        out.append("R" + files["b"] + files["c"] + files["d"] + "K" + files["f"] + files["g"] + "R")
    assert len(out) == 18
    return out


def chess324_fens():
    arrays = chess324_arrays()
    return [f"{b.lower()}/pppppppp/8/8/8/8/PPPPPPPP/{w} w KQkq - 0 1" for w in arrays for b in arrays]


def chess324_synth_book() -> np.ndarray:
    """'chess324-synth': every Chess324 start array followed by a seeded 8-ply random playout,
    keeping the positions after plies 0, 2, 4, 6, 8 -> 1 620 positions (fewer if a game ends)."""
    chunks = []
    for index, fen in enumerate(chess324_fens()):
        chunks.append(oracle.random_playout(oracle.parse(fen), 0xC3240000 + index, 8, 2))
    return np.ascontiguousarray(np.concatenate(chunks))


def playout_positions(target: int, seed: int = 0x5EED0000) -> np.ndarray:
    """Config #5: positions of seeded random playouts from the start position (even games) or a
    Chess324 array (odd games), every position recorded, games capped at 200 plies.  `seed` is the
    documented one of SURVEY.md 8(d) unless a tool asks for other games."""
    fens = chess324_fens()
    chunks, total, g = [], 0, 0
    start = oracle.starting()
    while total < target:
        root = start if g % 2 == 0 else oracle.parse(fens[(g // 2) % 324])
        arr = oracle.random_playout(root, seed + g, 200)
        chunks.append(arr)
        total += len(arr)
        g += 1
    return np.ascontiguousarray(np.concatenate(chunks)[:target])

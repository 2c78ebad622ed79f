"""A second opinion for the oracle (the reference uses shakmaty for this, tests/chess.rs:533-578 and
fuzz/fuzz_targets/generate_moves.rs:25-40): tests/native/brute.c, a mailbox generator that shares nothing with the oracle,
(1) is itself pinned by the reference's sorted move-list vectors, then (2) agrees with the oracle's move SETS and in_check
on more than a million positions: the 1 M seeded playout positions of BASELINE config 5, the Chess324 stand-in book and
the first plies of the reference's perft trees.  The GPU lists are compared with it in tests/test_gpu_parity.py."""
import ctypes as C
import time

import numpy as np
import pytest

import brute
import oracle
import workloads


def test_brute_force_generator_reproduces_the_reference_move_list_vectors(vectors):
    for case in vectors["move_lists"]:
        got = sorted(oracle.move_to_uci(m) for m in brute.legal_moves(oracle.parse(case["fen"])))
        assert got == case["moves"], case["source"]
    for case in vectors["move_counts"]:
        assert len(brute.legal_moves(oracle.parse(case["fen"]))) == case["count"], case


def test_brute_force_perft_of_the_small_reference_vectors(vectors):
    """The brute-force generator walks the reference's perft vectors by itself (make_move borrowed from the oracle)."""
    def perft(p, depth):
        moves = brute.legal_moves(p)
        if depth == 1:
            return len(moves)
        total = 0
        for m in moves:
            c = oracle.copy(p)
            oracle.make_move(c, m)
            total += perft(c, depth - 1)
        return total
    checked = 0
    for case in vectors["perft"]:
        if 0 < case["depth"] and case["nodes"] <= 20000:
            assert perft(oracle.parse(case["fen"]), case["depth"]) == case["nodes"], case
            checked += 1
    assert checked >= 20


def agree(positions):
    offsets, moves = oracle.generate_moves_batch(positions)
    checks = np.array([oracle.in_check(oracle.Position.from_buffer_copy(positions[i].tobytes())) for i in range(0, len(positions), 50)], dtype=np.uint8)
    bad, first, total = brute.compare_batch(positions, offsets, moves)
    assert bad == 0, oracle.to_fen(oracle.Position.from_buffer_copy(positions[first].tobytes()))
    assert total == len(moves)
    sample = np.ascontiguousarray(positions[::50])
    s_off, s_moves = oracle.generate_moves_batch(sample)
    bad, first, _ = brute.compare_batch(sample, s_off, s_moves, checks)
    assert bad == 0, ("in_check", oracle.to_fen(oracle.Position.from_buffer_copy(sample[first].tobytes())))
    return total


def test_oracle_agrees_with_the_brute_force_generator_on_a_million_positions():
    t0 = time.time()
    positions = workloads.playout_positions(1_000_000)
    assert agree(positions) == 27_639_035
    book = workloads.chess324_synth_book()
    assert agree(book) > 30_000
    other = workloads.playout_positions(60_000, seed=0x7A110000)      # games the fixtures never saw
    agree(other)
    assert time.time() - t0 < 240


def test_oracle_agrees_on_the_first_plies_of_the_reference_trees(vectors):
    """Checks, promotions, en passant, castling through attacked squares: the special positions of tests/chess.rs."""
    frontier = [oracle.parse(c["fen"]) for c in vectors["perft"] + vectors["bench_perft"] + vectors["move_lists"]]
    for _ in range(2):
        arr = oracle.positions_array(frontier)
        agree(arr)
        nxt = []
        for p in frontier[:400]:
            for m in oracle.generate_moves(p):
                c = oracle.copy(p)
                oracle.make_move(c, m)
                nxt.append(c)
        frontier = nxt[:: max(1, len(nxt) // 20000)]
    agree(oracle.positions_array(frontier))


def test_real_reference_inputs_when_present():
    """The reference's own book and 100 000-position file are absent from the mount (.MISSING_LARGE_BLOBS); when a mount
    has them they are read line by line with from_fen (tests/chess.rs:556-557) and put through the same comparison."""
    from pathlib import Path
    found = [p for p in (Path("/root/reference/books/Chess324_xxl_big_+090_+119.epd"), Path("/root/reference/tests/data/positions.fen")) if p.exists()]
    if not found:
        pytest.skip("the reference's book and positions.fen are not in this mount")
    for path in found:
        positions = [oracle.parse(line) for line in path.read_text().splitlines() if line.strip()]
        agree(oracle.positions_array(positions))

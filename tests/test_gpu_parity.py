"""GPU parity tests proper: the CUDA path, called through the C ABI
(include/pabi_cuda.h via pabi_b200), against the CPU oracle and the reference's
golden vectors (tests/golden/).  Integer work: everything is compared bit-exact.

Run on the B200 box with `python -m pytest tests -m gpu`.
"""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"


@pytest.fixture(scope="module")
def pb():
    import pabi_b200
    return pabi_b200


@pytest.fixture(scope="module")
def engine(pb):
    e = pb.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def small_engine(pb):
    """A context whose frontier buffers hold only 20 000 records: forces the chunked path."""
    e = pb.Engine(0, frontier_capacity=20000)
    yield e
    e.close()


@pytest.fixture(scope="module")
def playout_positions():
    """Seeded random-playout positions (SURVEY.md 8d, config #5 recipe), ~30k."""
    chunks = []
    start = oracle.starting()
    for g in range(400):
        chunks.append(oracle.random_playout(start, 0x5EED0000 + g, 200))
    return np.concatenate(chunks)


def fen_pos(pb, fen):
    return pb.Position.starting() if fen == "startpos" else pb.Position.try_from(fen)


# ---- perft: the reference's known answers (tests/chess.rs:622-818, benches/chess.rs:52-86) ----

def test_reference_perft_vectors(pb, engine, vectors):
    checked = 0
    for case in vectors["perft"] + vectors["bench_perft"]:
        if case["nodes"] > 2_000_000_000:
            continue
        assert engine.perft(fen_pos(pb, case["fen"]), case["depth"]) == case["nodes"], case
        checked += 1
    assert checked >= 70


def test_startpos_perft_7(pb, engine):
    assert engine.perft(pb.Position.starting(), 7) == 3_195_901_860


def test_reference_deep_vector(pb, engine, vectors):
    """tests/chess.rs:813-818: depth 7 -> 14 794 751 816."""
    deep = [c for c in vectors["perft"] if c["nodes"] > 2_000_000_000]
    assert deep
    for case in deep:
        assert engine.perft(fen_pos(pb, case["fen"]), case["depth"]) == case["nodes"], case


def test_perft_depth_0_1_2(pb, engine):
    p = pb.Position.from_fen(KIWIPETE)
    assert [engine.perft(p, d) for d in range(4)] == [1, 48, 2039, 97862]
    # no legal moves: checkmate and stalemate
    mate = pb.Position.from_fen("R5k1/5ppp/8/8/8/8/8/6K1 b - - 0 1")
    stale = pb.Position.from_fen("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")
    for p in (mate, stale):
        assert [engine.perft(p, d) for d in range(4)] == [1, 0, 0, 0]


def test_hashed_perft_is_exact(pb, engine, small_engine, vectors):
    """Transposition merging changes the work, never the count: every reference vector, also with chunked frontiers."""
    for case in vectors["perft"] + vectors["bench_perft"]:
        assert engine.perft_hashed(fen_pos(pb, case["fen"]), case["depth"]) == case["nodes"], case
        if case["nodes"] < 200_000_000:
            assert small_engine.perft_hashed(fen_pos(pb, case["fen"]), case["depth"]) == case["nodes"], case
    assert engine.perft_hashed(pb.Position.starting(), 8) == 84_998_978_956
    assert engine.last_timing.merged_records > 0
    for d in range(4):
        assert engine.perft_hashed(pb.Position.from_fen(KIWIPETE), d) == [1, 48, 2039, 97862][d]


def test_alternative_kernel_geometries_give_the_same_counts(pb, monkeypatch):
    """PABI_TILE_VARIANT selects other tile geometries of K2 (1..5), the warp-synchronous K2 (8) and the
    block-level K3 (9); they exist for timing comparisons and must all be exact."""
    p = pb.Position.from_fen(KIWIPETE)
    batch = oracle.random_playout(oracle.starting(), 0x5EED0003, 120)
    want_offsets, want_moves = oracle.generate_moves_batch(batch)
    for variant in ("1", "2", "3", "4", "5", "8", "9"):
        monkeypatch.setenv("PABI_TILE_VARIANT", variant)
        e = pb.Engine(0)
        assert e.perft(p, 5) == 193690690, variant
        assert e.perft(pb.Position.starting(), 6) == 119060324, variant
        offsets, moves = e.generate_moves_batch(batch)
        assert np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves), variant
        e.close()


def test_chunked_search_gives_the_same_counts(pb, small_engine, vectors):
    for case in vectors["perft"]:
        if 100_000 < case["nodes"] < 200_000_000:
            assert small_engine.perft(fen_pos(pb, case["fen"]), case["depth"]) == case["nodes"], case


def test_perft_batch_matches_oracle_per_root(engine, playout_positions):
    batch = np.ascontiguousarray(playout_positions[::37][:600])
    for depth in (1, 2, 3):
        got = engine.perft_batch(batch, depth)
        want = [oracle.perft(oracle.Position.from_buffer_copy(batch[i].tobytes()), depth) for i in range(len(batch))]
        assert got.tolist() == want, depth
    # few roots, deeper: several plies run inside the one-block root kernel, which has to carry the root indices
    few = np.ascontiguousarray(batch[:9])
    got = engine.perft_batch(few, 4)
    assert got.tolist() == [oracle.perft(oracle.Position.from_buffer_copy(few[i].tobytes()), 4) for i in range(len(few))]


def test_perft_batch_chunked(small_engine, playout_positions):
    batch = np.ascontiguousarray(playout_positions[::53][:300])
    got = small_engine.perft_batch(batch, 4)
    want = [oracle.perft(oracle.Position.from_buffer_copy(batch[i].tobytes()), 4) for i in range(len(batch))]
    assert got.tolist() == want


def test_shards_sum_to_the_whole(pb, engine):
    p = pb.Position.from_fen(KIWIPETE)
    for depth, split in ((4, 1), (4, 2), (5, 2), (3, 5), (1, 2), (2, 1), (3, 2), (4, 0)):
        whole = engine.perft(p, depth)
        for count in (1, 2, 3, 8):
            assert sum(engine.perft_shard(p, depth, split, i, count) for i in range(count)) == whole


def test_perft_multi_contexts_in_one_process(pb, engine, small_engine):
    """pabi_cuda_perft_multi: several contexts (here two on the same device, one of them with a tiny
    frontier), one host thread each, host-side sum."""
    import torch
    p = pb.Position.from_fen(KIWIPETE)
    nodes, ms = pb.perft_multi([engine, small_engine], p, 4, split_ply=2)
    assert nodes == 4085603 and len(ms) == 2
    if torch.cuda.device_count() >= 2:
        other = pb.Engine(1)
        nodes, _ = pb.perft_multi([engine, other], pb.Position.starting(), 6)
        assert nodes == 119060324
        other.close()


def test_divide(pb, engine):
    p = pb.Position.from_fen(KIWIPETE)
    op = oracle.parse(KIWIPETE)
    got = engine.divide(p, 3)
    want = []
    for m in oracle.generate_moves(op):
        c = oracle.copy(op)
        oracle.make_move(c, m)
        want.append((m, oracle.perft(c, 2)))
    assert [(m.raw, n) for m, n in got] == want
    assert sum(n for _, n in got) == 97862


# ---- move lists: golden vectors (tests/chess.rs:219-472) and order vs the oracle ----

def test_reference_move_list_vectors(pb, engine, vectors):
    for case in vectors["move_lists"]:
        p = fen_pos(pb, case["fen"])
        assert sorted(str(m) for m in p.generate_moves(engine)) == case["moves"], case["source"]
    for case in vectors["move_counts"]:
        assert len(fen_pos(pb, case["fen"]).generate_moves(engine)) == case["count"], case["source"]


def test_move_lists_bit_exact_on_playout_positions(engine, playout_positions):
    offsets, moves = engine.generate_moves_batch(playout_positions)
    want_offsets, want_moves = oracle.generate_moves_batch(playout_positions)
    assert np.array_equal(offsets, want_offsets)
    assert np.array_equal(moves, want_moves)
    counts = engine.count_moves_batch(playout_positions)
    assert np.array_equal(counts, np.diff(want_offsets))


def test_page_locked_host_arrays(pb, engine, playout_positions):
    """pabi_cuda_host_alloc: the batched calls give the same answers from and into page-locked arrays."""
    n = 5000
    src = np.ascontiguousarray(playout_positions[:n])
    pinned = pb.pinned_empty(n, src.dtype)
    pinned[:] = src
    offsets_out, moves_out = pb.pinned_empty(n + 1, np.uint32), pb.pinned_empty(64 * n, np.uint16)
    offsets, moves = engine.generate_moves_batch(pinned, offsets_out, moves_out)
    want_offsets, want_moves = oracle.generate_moves_batch(src)
    assert np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves)
    assert np.shares_memory(moves, moves_out) and np.shares_memory(offsets, offsets_out)
    assert np.array_equal(engine.count_moves_batch(pinned), np.diff(want_offsets))


def test_empty_and_tiny_batches(engine, playout_positions):
    offsets, moves = engine.generate_moves_batch(playout_positions[:0])
    assert offsets.tolist() == [0] and len(moves) == 0
    assert len(engine.perft_batch(playout_positions[:0], 3)) == 0
    offsets, moves = engine.generate_moves_batch(playout_positions[:1])
    assert offsets.tolist() == [0, 20]


def test_in_check(engine, playout_positions):
    got = engine.in_check_batch(playout_positions[:5000])
    want = [oracle.in_check(oracle.Position.from_buffer_copy(playout_positions[i].tobytes())) for i in range(5000)]
    assert got.tolist() == want


def test_attack_info(pb, engine, vectors, playout_positions):
    """AttackInfo pictures of attacks.rs:555-695 and the oracle on playout positions, five fields each."""
    for case in vectors["attacks"]["attack_info"]:
        got = fen_pos(pb, case["fen"]).attack_info(engine)
        for field, bits in case["fields"].items():
            assert got[field] == int(bits), (case["source"], field)
    batch = np.ascontiguousarray(playout_positions[:6000])
    got = engine.attack_info_batch(batch)
    for i in range(len(batch)):
        ai = oracle.attack_info(oracle.Position.from_buffer_copy(batch[i].tobytes()))
        assert got[i].tolist() == [ai.attacks, ai.checkers, ai.pins, ai.xrays, ai.safe_king_squares]


# ---- make_move: children of a ply, field for field (tests/chess.rs:474-531,580-620) ----

def test_reference_make_move_vectors(pb, engine, vectors):
    for case in vectors["make_move"]:
        p = fen_pos(pb, case["start"])
        for step in case["steps"]:
            if "move" in step:
                p.make_move(pb.Move.from_uci(step["move"]), engine)
            else:
                assert str(p) == step["fen"], case["test"]


def test_children_equal_oracle_children(engine, playout_positions):
    batch = np.ascontiguousarray(playout_positions[::11][:2500])
    offsets, children = engine.expand_batch(batch)
    want_offsets, want_moves = oracle.generate_moves_batch(batch)
    assert np.array_equal(offsets, want_offsets)
    k = 0
    for i in range(len(batch)):
        parent = oracle.Position.from_buffer_copy(batch[i].tobytes())
        for m in want_moves[want_offsets[i]:want_offsets[i + 1]]:
            c = oracle.copy(parent)
            oracle.make_move(c, int(m))  # hash included: the device maintains Position::hash exactly as make_move does
            assert bytes(c) == children[k].tobytes(), (oracle.to_fen(parent), oracle.move_to_uci(int(m)))
            k += 1
    assert k == len(children)


def test_random_playouts_on_device_equal_the_oracle_driver(pb, engine):
    """The datagen-style position source run on the device: every recorded position of every game equals
    the CPU driver's (oracle_random_playout), clocks and en passant squares included."""
    import workloads
    fens = workloads.chess324_fens()
    starts, seeds = [], []
    for g in range(600):
        starts.append(oracle.starting() if g % 2 == 0 else oracle.parse(fens[(g // 2) % 324]))
        seeds.append(0x5EED0000 + g)
    start_arr = oracle.positions_array(starts)
    for max_plies, stride in ((200, 1), (40, 3), (0, 1)):
        games = engine.random_playouts(start_arr, seeds, max_plies, stride)
        total = 0
        for g, got in enumerate(games):
            want = oracle.random_playout(starts[g], seeds[g], max_plies, stride)
            assert got.tobytes() == want.tobytes(), (g, max_plies, stride)
            total += len(got)
        assert total >= 600


def test_books_parsed_on_the_device(pb, engine, vectors, playout_positions):
    """pabi_cuda_parse_batch: every line gets the host parser's verdict and, when accepted, its position."""
    import random
    from test_host_api import REJECTED
    rng = random.Random(5)
    lines = list(vectors["fen"]["roundtrip"]) + list(vectors["fen"]["try_from_ok"]) + list(vectors["fen"]["try_from_err"])
    lines += REJECTED + [c["fen"] for c in vectors["fen"]["errors"]]
    for i in range(0, 20000, 7):
        fen = oracle.to_fen(oracle.Position.from_buffer_copy(playout_positions[i].tobytes()))
        lines.append(fen if i % 3 else "epd " + " ".join(fen.split()[:4]) + "\n")
        chars = list(fen)
        chars[rng.randrange(len(chars))] = rng.choice("pnbrqkPNBRQK12345678/ wb-KQkqabcdefgh0")
        lines.append("".join(chars))
    positions, status = engine.parse_batch(lines)
    accepted = 0
    for i, line in enumerate(lines):
        try:
            want = pb.Position.try_from(line)
        except ValueError:
            assert status[i] < 0, line
            continue
        assert status[i] == 0, line
        assert positions[i].tobytes() == bytes(want.raw), line
        accepted += 1
    assert 3000 < accepted < len(lines)


# ---- Game environment (src/chess/game.rs:21-111 without the Syzygy branch) ----

def test_reference_game_tests(pb, engine):
    """game.rs:143-227: repetition, stalemate, checkmate, fifty-move rule -- four games in one batch."""
    fens = ["rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
            "3b2qk/p6p/1p3Q1P/8/8/n7/PP6/K7 b - - 3 2",
            "3b3k/p5qp/1p3Q1P/8/8/n7/PP6/K7 w - - 4 3",
            "8/5k2/3p4/1p1Pp2p/pP2Pp1P/P4P1K/8/8 b - - 99 50"]
    games = pb.Games(engine, [pb.Position.from_fen(f) for f in fens], history_cap=16)
    assert games.result() == [None, None, None, None]
    games.apply([pb.Move.from_uci("g1f3"), pb.Move.from_uci("d8f6"), pb.Move.from_uci("f6g7"), pb.Move.from_uci("f7f6")])
    assert games.result() == [None, "draw", "win", "draw"]
    offsets, moves = games.actions()
    assert offsets[2] - offsets[1] == 0 and offsets[3] - offsets[2] == 0      # stalemate and checkmate: no actions
    for i, u in enumerate(["g8f6", "f3g1", "f6g8", "g1f3", "g8f6", "f3g1"]):
        games.apply([pb.Move.from_uci(u), None, None, None])
        assert games.result()[0] is None, i
    games.apply([pb.Move.from_uci("f6g8"), None, None, None])
    assert games.result() == ["draw", "draw", "win", "draw"]                  # threefold repetition
    assert str(pb.Position.from_record(games.positions()[0])).startswith("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -")
    with pytest.raises(pb.PabiCudaError):                                     # history_cap = 16 actions
        for _ in range(12):
            games.apply([pb.Move.from_uci("g1f3"), None, None, None])
            games.apply([pb.Move.from_uci("g8f6"), None, None, None])
            games.apply([pb.Move.from_uci("f3g1"), None, None, None])
            games.apply([pb.Move.from_uci("f6g8"), None, None, None])
    games.close()


def test_repetition_semantics_follow_the_reference(pb, engine):
    """tests/test_oracle_game.py::test_repetition_key_ignores_the_side_to_move_like_the_reference on the device."""
    from test_oracle_game import TRIANGULATION
    games = pb.Games(engine, [pb.Position.from_fen("8/8/8/3k4/8/3K4/8/8 w - - 0 1")], history_cap=32)
    for u in TRIANGULATION[:-1]:
        games.apply([pb.Move.from_uci(u)])
        assert games.result() == [None]
    games.apply([pb.Move.from_uci(TRIANGULATION[-1])])
    assert games.result() == ["draw"]
    games.close()


def test_random_self_play_equals_the_oracle_game(pb, engine):
    """Whole games on the device: result, length and final position of every game equal the CPU `Game`."""
    import workloads
    fens = workloads.chess324_fens()
    roots = [oracle.starting() if g % 3 else oracle.parse(fens[g % 324]) for g in range(1500)]
    roots += [oracle.from_fen("8/5k2/3p4/1p1Pp2p/pP2Pp1P/P4P1K/8/8 b - - 90 50"), oracle.from_fen("4k3/8/8/8/8/8/8/R3K3 w Q - 0 1")] * 20
    seeds = [0x6A3E0000 + g for g in range(len(roots))]
    games = pb.Games(engine, oracle.positions_array(roots), history_cap=700)
    results, plies = games.play_random(seeds, 600)
    finals = games.positions()
    tally = {}
    for g, root in enumerate(roots):
        og = oracle.Game(root)
        r, n = og.play_random(seeds[g], 600)
        assert (results[g], int(plies[g])) == (r, n), g
        want = og.position()
        assert bytes(want) == finals[g].tobytes(), g  # all 112 bytes, Position::hash included
        tally[r] = tally.get(r, 0) + 1
    assert tally.get("draw", 0) > 0 and (tally.get("win", 0) + tally.get("loss", 0)) > 0
    assert games.result() == results                                          # Game::result agrees after the fact
    games.close()


# ---- text API through the same library (tests/chess.rs:33-181) ----

def test_fen_round_trips_and_errors(pb, vectors):
    fen = vectors["fen"]
    for text in fen["roundtrip"]:
        assert str(pb.Position.from_fen(text)).startswith(text)


# ---- BASELINE.json configs #3 and #5 at full size ----

def test_chess324_book_perft4_bit_exact(engine):
    """configs[2]: batched perft(4) over every position of the Chess324 book.  The book file is
    absent from the reference mount; the documented stand-in (tests/workloads.py) is used and
    every count is checked against the oracle, castling included."""
    import workloads
    book = workloads.chess324_synth_book()
    got = engine.perft_batch(book, 4)
    threads = min(16, __import__("os").cpu_count() or 1)
    want = [oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), 4, threads) for i in range(len(book))]
    assert got.tolist() == want
    assert sum(want) > 500_000_000


def colour_flipped(pb, p):
    """The same position seen from the other side: boards mirrored top to bottom, colours, castling rights and the
    side to move swapped.  Chess is symmetric under this map, so every perft count is preserved."""
    def flip(b):
        return int.from_bytes(int(b).to_bytes(8, "little"), "big")  # rank 1 <-> rank 8
    q = pb.Position()
    r, o = q.raw, p.raw
    for k in range(6):
        r.white[k], r.black[k] = flip(o.black[k]), flip(o.white[k])
    r.castling = ((o.castling & 3) << 2) | (o.castling >> 2)
    r.side_to_move = o.side_to_move ^ 1
    r.halfmove_clock, r.fullmove_counter = o.halfmove_clock, o.fullmove_counter
    r.en_passant_square = 64 if o.en_passant_square >= 64 else o.en_passant_square ^ 56
    return q


def test_colour_symmetry_at_full_size(pb, engine, vectors):
    """A size-independent property at the sizes the oracle cannot reach in seconds: the colour-flipped position has
    the same perft count (the kernels take different code paths for it: the K2 parents are mirrored or not)."""
    big = [c for c in vectors["perft"] + vectors["bench_perft"] if 50_000_000 < c["nodes"] < 1_000_000_000]
    assert len(big) >= 5
    for case in big:
        p = fen_pos(pb, case["fen"])
        assert engine.perft(colour_flipped(pb, p), case["depth"]) == case["nodes"], case
    # and per root over a batch
    flipped = [colour_flipped(pb, fen_pos(pb, c["fen"])) for c in big[:4]]
    assert engine.perft_batch(flipped, 4).tolist() == engine.perft_batch([fen_pos(pb, c["fen"]) for c in big[:4]], 4).tolist()


def test_dense_position_large_tiles(pb, engine):
    """Ten queens a side: ~120 moves per position, almost none of them quiet, so a 512-parent tile of K2 holds about
    62 000 pool entries of one class -- the per-class prefix sums of a tile must not be 16-bit.  Counts from the oracle."""
    p = pb.Position.try_from("knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1")
    assert engine.perft(p, 3) == 1_664_938
    assert engine.perft(p, 4) == 169_711_264
    assert engine.perft(p, 5) == 19_678_195_927
    # the same family through the list kernels: a warp's 32 positions hold ~3 800 moves, twice its shared-memory pool
    _, ply1 = engine.expand_batch([p])
    _, ply2 = engine.expand_batch(ply1)
    offsets, moves = engine.generate_moves_batch(ply2)
    want_offsets, want_moves = oracle.generate_moves_batch(np.ascontiguousarray(ply2))
    assert len(ply2) == 13_774 and np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves)
    assert engine.perft_batch(ply2[:700], 2).tolist() == [
        oracle.perft(oracle.Position.from_buffer_copy(ply2[i].tobytes()), 2) for i in range(700)]


def test_command_line(pb):
    """python -m pabi_b200: the OpenBench line the reference's `openbench()` is meant to print (src/engine/mod.rs:176-190,
    tests/integration.rs:23-34: "<nodes> nodes <nps> nps") over the positions of benches/chess.rs:52-86, and perft divide."""
    import re
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, "-m", "pabi_b200", "bench"], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1000:]
    m = re.fullmatch(r"(\d+) nodes (\d+) nps", out.stdout.strip().splitlines()[-1])
    assert m and int(m.group(1)) == 1_461_874_231 and int(m.group(2)) > 0
    out = subprocess.run([sys.executable, "-m", "pabi_b200", "perft", "--depth", "3", "--divide", "--fen", KIWIPETE], cwd=root,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1000:]
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "Nodes searched: 97862" and len(lines) == 48 + 2 and lines[0].startswith("e1d1: ")


def test_one_million_move_lists_bit_exact(engine):
    """configs[4]: ordered move lists for 1M random-playout positions, bit-exact vs the oracle."""
    import workloads
    positions = workloads.playout_positions(1_000_000)
    offsets, moves = engine.generate_moves_batch(positions)
    want_offsets, want_moves = oracle.generate_moves_batch(positions)
    assert np.array_equal(offsets, want_offsets)
    assert np.array_equal(moves, want_moves)
    assert len(moves) > 20_000_000
    # the same lists against the independent brute-force generator (sorted sets, in_check too), as the reference
    # does against shakmaty (tests/chess.rs:555-578)
    import brute
    checks = engine.in_check_batch(positions)
    bad, first, total = brute.compare_batch(positions, offsets, moves, checks)
    assert bad == 0 and total == len(moves), oracle.to_fen(oracle.Position.from_buffer_copy(positions[first].tobytes()))
    # and the two-pass path of round 1 (count, scan, emit) is still exact
    two_pass = np.ascontiguousarray(positions[:200_000])
    import os
    os.environ["PABI_LIST_TWO_PASS"] = "1"
    try:
        import pabi_b200
        e2 = pabi_b200.Engine(0)
        o2, m2 = e2.generate_moves_batch(two_pass)
        e2.close()
    finally:
        del os.environ["PABI_LIST_TWO_PASS"]
    assert np.array_equal(o2, want_offsets[:200_001]) and np.array_equal(m2, want_moves[:want_offsets[200_000]])


def test_startpos_perft_8(pb, engine):
    """configs[3]: 84 998 978 956 nodes."""
    assert engine.perft(pb.Position.starting(), 8) == 84_998_978_956


# ---- round 2: Position::hash on the device, per-root depths, the C consumer, refused records ----

def test_reference_hash_properties_with_make_move_on_the_device(pb, engine):
    """tests/chess.rs:820-868 (repetition_hash, castling_hash): the hash is maintained by the device make_move."""
    p = pb.Position.try_from("8/5k2/6p1/8/8/8/1p3P2/5K2 w - - 0 1")
    initial = p.hash()
    assert initial == p.compute_hash() != 0
    for uci in ("f1e2", "f7f6", "e2f1"):
        p.make_move(pb.Move.from_uci(uci), engine)
        assert p.hash() != initial
    p.make_move(pb.Move.from_uci("f6f7"), engine)
    assert str(p) == "8/5k2/6p1/8/8/8/1p3P2/5K2 w - - 4 3" and p.hash() == initial
    p = pb.Position.try_from("rnbqk1nr/p3bppp/1p2p3/2ppP3/3P4/P7/1PP1NPPP/R1BQKBNR w KQkq - 0 7")
    initial = p.hash()
    for uci in ("e1d2", "e8d7", "d2e1", "d7e8"):
        p.make_move(pb.Move.from_uci(uci), engine)
    assert str(p) == "rnbqk1nr/p3bppp/1p2p3/2ppP3/3P4/P7/1PP1NPPP/R1BQKBNR w - - 4 9" and p.hash() != initial


def test_hash_of_every_child_equals_the_oracles(pb, engine, vectors):
    """Device `hash` == oracle `hash` on every child of the reference trees (two plies below every perft vector's
    root, all 112 bytes compared), and on every line parsed on the device."""
    roots = [oracle.parse(c["fen"]) for c in vectors["perft"] + vectors["bench_perft"]]
    level = oracle.positions_array(roots)
    for _ in range(2):
        offsets, children = engine.expand_batch(level)
        want_offsets, want_moves = oracle.generate_moves_batch(level)
        assert np.array_equal(offsets, want_offsets)
        want = np.zeros(len(children), dtype=level.dtype)
        k = 0
        for i in range(len(level)):
            parent = oracle.Position.from_buffer_copy(level[i].tobytes())
            for m in want_moves[want_offsets[i]:want_offsets[i + 1]]:
                c = oracle.copy(parent)
                oracle.make_move(c, int(m))
                C.memmove(want.ctypes.data + 112 * k, C.byref(c), 112)
                k += 1
        assert k == len(children) and children.tobytes() == want.tobytes()
        assert children["hash"].all()  # the suite repeats roots at several depths, so equal hashes are expected; zero is not
        level = np.ascontiguousarray(children[:: max(1, len(children) // 3000)])
    lines = [oracle.to_fen(oracle.Position.from_buffer_copy(level[i].tobytes())) for i in range(0, len(level), 3)]
    positions, status = engine.parse_batch(lines)
    assert not status.any()
    for line, got in zip(lines, positions):
        assert got.tobytes() == bytes(oracle.from_fen(line))


def test_perft_suite_as_one_batch_with_per_root_depths(pb, engine, small_engine, vectors):
    """the whole reference suite through one call (`pabi_cuda_perft_batch_depths`)."""
    cases = [c for c in vectors["perft"] + vectors["bench_perft"] if c["nodes"] < 2_000_000_000]
    roots = oracle.positions_array([oracle.parse(c["fen"]) for c in cases])
    depths = [c["depth"] for c in cases]
    assert engine.perft_batch_depths(roots, depths).tolist() == [c["nodes"] for c in cases]
    light = [i for i, c in enumerate(cases) if c["nodes"] < 50_000_000]
    got = small_engine.perft_batch_depths(np.ascontiguousarray(roots[light]), [depths[i] for i in light])
    assert got.tolist() == [cases[i]["nodes"] for i in light]
    assert engine.perft_batch_depths(roots[:0], []).tolist() == []


def test_records_that_are_not_positions_are_refused(pb, engine):
    bad = np.zeros(3, dtype=pb.POSITION_DTYPE)
    bad[0] = np.frombuffer(bytes(pb.Position.starting().raw), dtype=pb.POSITION_DTYPE)[0]
    bad[2] = bad[0]
    bad[2]["white"][5] |= 1 << 60                       # a white pawn on the eighth rank
    for call in (lambda: engine.perft_batch(bad, 3), lambda: engine.generate_moves_batch(bad), lambda: engine.count_moves_batch(bad),
                 lambda: pb.Games(engine, bad)):
        with pytest.raises(pb.PabiCudaError) as e:
            call()
        assert e.value.status == -5 and "not" in str(e.value)
    assert engine.perft(pb.Position.starting(), 4) == 197281  # the context is unharmed


def test_c_consumer_of_the_header_runs_a_perft(tmp_path):
    """tests/native/abi_smoke.c linked against the library: struct layouts asserted with offsetof, one perft and one
    per-root-depth batch through include/pabi_cuda.h from plain C."""
    import subprocess
    from pathlib import Path
    from pabi_b200 import _native
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", str(root / "include"), str(root / "tests/native/abi_smoke.c"),
                    "-o", str(exe), "-L", str(_native.PKG_DIR), "-lpabi_b200", f"-Wl,-rpath,{_native.PKG_DIR}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "device perft ok" in out.stdout, out.stdout + out.stderr


def test_games_on_one_device_while_another_context_is_current(pb, engine):
    """every entry point selects its context's device before it allocates.  With two devices, calls on
    Engine(1) are interleaved with a Games object that lives on Engine(0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    other = pb.Engine(1)
    games = pb.Games(engine, oracle.positions_array([oracle.starting()] * 64), history_cap=64)
    assert other.perft(pb.Position.starting(), 4) == 197281           # leaves device 1 current on this thread
    offsets, moves = games.actions()
    assert offsets[-1] == 64 * 20
    assert other.perft(pb.Position.starting(), 3) == 8902
    games.apply([int(moves[offsets[g]]) for g in range(64)])
    assert games.result() == [None] * 64
    games.close()
    other.close()


def test_device_side_workload_generators_reproduce_the_oracle_workloads(pb, engine):
    """pabi_b200/workloads.py builds the stand-in book and the 1 M playout positions with the device playout kernel; both are
    byte-identical (Position::hash included) to the oracle-built workloads whose digests tests/golden/workload_fixtures.json holds."""
    import hashlib
    import json
    from pathlib import Path
    from pabi_b200 import workloads as W
    fx = json.loads((Path(__file__).resolve().parent / "golden" / "workload_fixtures.json").read_text())
    book = W.chess324_synth_book(engine)
    assert len(book) == 1620 and hashlib.sha256(book.tobytes()).hexdigest() == fx["book"]["sha256"]
    assert engine.perft_batch(book, 4).tolist() == fx["book"]["perft"]
    positions = W.playout_positions(engine, 1_000_000)
    assert hashlib.sha256(positions.tobytes()).hexdigest() == fx["playouts"]["sha256"]


def test_no_index_left_its_array_in_a_checking_build(pb, engine, small_engine):
    """With the -DPABI_BOUNDS library in place (tools/bounds_suite.sh) every record and pool index of every kernel launched
    by the tests above was range-checked; the shipped build answers -1 (no checks compiled in).  Runs last in this file."""
    for e in (engine, small_engine):
        assert e.bounds_violation() in (0, -1), f"out-of-range index at chess.cuh/kernels.cuh line {e.bounds_violation()}"


def test_speculative_plies_fall_back_to_the_exact_path_when_a_buffer_overflows(pb):
    """Small plies are expanded without the host knowing their size, on the estimate of 32 children per position; a ply that
    does not fit after all raises the overflow flag and the call is redone on the exact, chunking path.  A 4 096-record
    frontier and a position with ~120 moves per node force exactly that."""
    dense = pb.Position.try_from("knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1")
    tiny = pb.Engine(0, frontier_capacity=4096)
    assert tiny.perft(dense, 3) == 1664938
    assert tiny.perft(dense, 4) == 169711264
    assert sum(tiny.perft_shard(dense, 4, 2, i, 3) for i in range(3)) == 169711264
    assert tiny.perft(pb.Position.starting(), 6) == 119060324
    tiny.close()


def test_shard_assignment_is_the_documented_hash_of_the_record(pb, engine, playout_positions):
    """pabi_cuda_perft_shard keeps the subtrees whose root record hashes to its shard (pabi_b200.distributed.shard_of restates
    the hash on the host).  With split_ply 0 the root itself is the subtree: exactly one shard returns its count."""
    from pabi_b200 import distributed as D
    for i in range(0, 3000, 97):
        p = pb.Position.from_record(playout_positions[i])
        whole = engine.perft(p, 2)
        for count in (2, 8):
            owner = D.shard_of(p.pack64(), count)
            got = [engine.perft_shard(p, 2, 0, s, count) for s in range(count)]
            assert got == [whole if s == owner else 0 for s in range(count)], (i, count)


def test_virtual_ply_and_materialised_ply_give_the_same_counts(pb, monkeypatch):
    """With three plies left the ply above K2 is not materialised (K1 writes 8-byte (parent, move) entries, K2 makes its
    parents itself); PABI_NO_VIRTUAL_PLY=1 keeps the round-1 behaviour.  Both must agree everywhere: white and black
    grandparents (the mirrored make_move), castling / en passant / promotion positions, per-root batches, and a frontier so
    small (4 096 records = 32 768 entries) that the virtual ply is cut into several chunks."""
    cases = [(KIWIPETE, 3, 97862), (KIWIPETE, 4, 4085603), (KIWIPETE, 5, 193690690),
             ("n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1", 5, 3605103), ("n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1", 6, 71179139),
             ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 6, 11030083),
             ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 5, 15833292)]
    roots = [pb.Position.try_from(fen) for fen, _, _ in cases]
    for env in ("0", "1"):
        monkeypatch.setenv("PABI_NO_VIRTUAL_PLY", env)
        e, tiny = pb.Engine(0), pb.Engine(0, frontier_capacity=4096)
        for (fen, depth, nodes), root in zip(cases, roots):
            assert e.perft(root, depth) == nodes, (env, fen, depth)
            if nodes < 20_000_000:
                assert tiny.perft(root, depth) == nodes, (env, "tiny", fen, depth)
        assert e.perft_batch(roots, 3).tolist() == tiny.perft_batch(roots, 3).tolist() == [pb.Engine.perft(e, r, 3) for r in roots]
        assert sum(e.perft_shard(roots[0], 5, 2, i, 3) for i in range(3)) == 193690690
        e.close(), tiny.close()


def test_move_lists_of_positions_with_more_moves_than_a_staging_column(pb, engine, vectors):
    """The fused list kernel stages at most 64 moves per position; a warp that holds a fatter position regenerates its lists
    through a sliding window.  Positions with up to 218 legal moves, alone, repeated, and mixed into ordinary batches of
    ragged sizes, must still come out in the reference's order."""
    fat = ["knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1",           # ~120 moves
           "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1",                  # 218 moves, the record
           "3Q4/1Q4Q1/4Q3/2Q4R/Q4Q2/3Q4/1Q4Rp/1K1BBNNk w - - 0 1",                  # 218 moves
           "n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1"]                              # promotions on both wings
    fat_positions = [oracle.parse(f) for f in fat]
    assert max(len(oracle.generate_moves(p)) for p in fat_positions) == 218
    ordinary = [oracle.parse(c["fen"]) for c in vectors["move_lists"]]
    for batch in (fat_positions, fat_positions * 40, ordinary + fat_positions, (ordinary * 3 + fat_positions[:2]) * 7, fat_positions[1:2] * 33):
        arr = oracle.positions_array(batch)
        offsets, moves = engine.generate_moves_batch(arr)
        want_offsets, want_moves = oracle.generate_moves_batch(arr)
        assert np.array_equal(offsets, want_offsets) and np.array_equal(moves, want_moves), len(batch)
        # and with an output array that is too small: the required size comes back, nothing is written past the end
        small = np.full(16, 0xFFFF, dtype=np.uint16)
        offsets2, moves2 = engine.generate_moves_batch(arr, None, small)
        assert np.array_equal(offsets2, want_offsets) and np.array_equal(moves2, want_moves) and (small == 0xFFFF).all()

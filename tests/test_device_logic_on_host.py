"""The product's device arithmetic (pabi_b200/csrc/chess.cuh, compiled for the
host by tests/native) against the oracle and the reference's golden vectors.
This is the no-GPU half of the parity evidence: the kernels inline exactly this
source."""
import ctypes as C

import numpy as np
import pytest

import hostcheck
import oracle

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"


def gen(p):
    out = (C.c_uint16 * 256)()
    n = hostcheck.lib().hostcheck_generate_moves(C.byref(p), out)
    return list(out[:n])


def test_tables_match_the_oracles():
    L = hostcheck.lib()
    for name, oname in (("knight", "knight_attacks"), ("king", "king_attacks"), ("rays", "rays")):
        n = C.c_size_t()
        ptr = L.hostcheck_table(name.encode(), C.byref(n))
        got = np.ctypeslib.as_array(ptr, shape=(n.value,))
        np.testing.assert_array_equal(got, oracle.table(oname), err_msg=name)
    # slider blocks keep pabi's masks, offsets and sizes (generated.rs:30-56); entries are
    # permuted inside a block, so compare as sorted multisets per square ...
    for rook, tname, oname in ((0, "bishop_attacks", "bishop"), (1, "rook_attacks", "rook")):
        n = C.c_size_t()
        ptr = L.hostcheck_table(tname.encode(), C.byref(n))
        got = np.ctypeslib.as_array(ptr, shape=(n.value,))
        want = oracle.table(f"{oname}_attacks")
        masks = oracle.table(f"{oname}_relevant_occupancies")
        offsets = oracle.table(f"{oname}_attack_offsets")
        assert len(got) == len(want)
        for s in range(64):
            mask, off, bits = C.c_uint64(), C.c_uint32(), C.c_uint32()
            L.hostcheck_slider_entry(rook, s, C.byref(mask), C.byref(off), C.byref(bits))
            assert mask.value == int(masks[s]) and off.value == int(offsets[s])
            assert bits.value == bin(mask.value).count("1")
            blk = slice(off.value, off.value + (1 << bits.value))
            assert set(np.unique(want[blk])) >= set(np.unique(got[blk])) - {0}


def test_slider_lookups_equal_pext_lookups_on_random_occupancies():
    rng = np.random.default_rng(7)
    L, O = hostcheck.lib(), oracle.lib()
    for _ in range(20000):
        occ = int(rng.integers(0, 2**63, dtype=np.uint64)) & int(rng.integers(0, 2**63, dtype=np.uint64))
        s = int(rng.integers(0, 64))
        assert L.hostcheck_bishop(s, occ) == O.oracle_bishop_attacks(s, occ)
        assert L.hostcheck_rook(s, occ) == O.oracle_rook_attacks(s, occ)
    # attacks.rs:224-302
    assert L.hostcheck_rook(28, 0x1000404825001002) == 0x101010102C101000
    assert L.hostcheck_bishop(28, 0x1000404825001002) == 0x0000402800284482


def test_reference_move_list_vectors(vectors):
    for case in vectors["move_lists"]:
        p = oracle.parse(case["fen"])
        moves = gen(p)
        assert sorted(oracle.move_to_uci(m) for m in moves) == case["moves"], case["source"]
        assert moves == oracle.generate_moves(p)  # and in the reference's order
        assert hostcheck.lib().hostcheck_count_moves(C.byref(p)) == len(moves)
    for case in vectors["move_counts"]:
        p = oracle.parse(case["fen"])
        assert hostcheck.lib().hostcheck_count_moves(C.byref(p)) == case["count"], case["source"]


def test_reference_make_move_vectors(vectors):
    for case in vectors["make_move"]:
        p = oracle.parse(case["start"])
        for step in case["steps"]:
            if "move" in step:
                hostcheck.lib().hostcheck_make_move(C.byref(p), oracle.move_from_uci(step["move"]))
            else:
                assert oracle.to_fen(p) == step["fen"], case["test"]


def test_reference_perft_vectors(vectors):
    done = 0
    for case in vectors["perft"]:
        if case["nodes"] > 3_000_000:
            continue
        p = oracle.parse(case["fen"])
        assert hostcheck.lib().hostcheck_perft(C.byref(p), case["depth"]) == case["nodes"], case
        done += 1
    assert done >= 50


def test_analysis_matches_attack_info(vectors):
    for fen in list(vectors["attacks"]) if isinstance(vectors["attacks"], dict) else []:
        pass
    fens = [KIWIPETE, "startpos", "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
            "r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1"]
    for fen in fens:
        p = oracle.parse(fen)
        out = (C.c_uint64 * 5)()
        hostcheck.lib().hostcheck_analysis(C.byref(p), out)
        ai = oracle.attack_info(p)
        # analyze() builds the enemy attack map only where move generation reads it (king steps, castling
        # walks), so `attacks` is compared through safe_king_squares; the full picture is attack_info()'s job
        assert out[1] == ai.checkers and out[2] == ai.pins and out[3] == ai.safe_king_squares


ROOTS = [
    ("startpos", 4),
    (KIWIPETE, 3),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 5),
    ("r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1", 4),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 4),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 3),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", 3),
    ("8/Pk6/8/8/8/8/6KP/8 w - - 0 1", 6),
    ("8/8/1p4k1/1P6/8/8/6K1/8 w - - 0 1", 6),
    ("8/5k2/6p1/8/8/8/1p3P2/5K2 w - - 0 1", 6),
    ("rnb1kbnr/pp1pp1pp/1qp2p2/8/Q1P5/N7/PP1PPPPP/1RB1KBNR b Kkq - 2 4", 3),
    # en passant with pins / discovered checks
    ("8/8/8/2k5/3Pp3/8/8/4K3 b - d3 0 1", 4),
    ("8/8/8/8/k2Pp2Q/8/8/3K4 b - d3 0 1", 4),
    ("3k4/3p4/8/K1P4r/8/8/8/8 b - - 0 1", 5),
    ("8/8/4k3/8/2p5/8/B2P2K1/8 w - - 0 1", 5),
    ("r3k2r/8/3Q4/8/8/5q2/8/R3K2R b KQkq - 0 1", 3),
    ("2kr3r/p1ppqpb1/bn2Qnp1/3PN3/1p2P3/2N5/PPPBBPPP/R3K2R b KQ - 3 2", 3),
    ("8/8/2k5/5q2/5n2/8/5K2/8 b - - 0 1", 4),
    ("4k3/1P6/8/8/8/8/K7/8 w - - 0 1", 5),
    ("n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1", 4),
    # castling rights that the board does not back (the reference trusts the bits,
    # position.rs:1236-1286, :632-678)
    ("4k3/8/8/8/8/8/8/4K2N w K - 0 1", 3),
    ("4k2r/8/8/8/8/8/8/4K2r w K - 0 1", 3),
    ("r3k3/8/8/8/8/8/8/n3K3 w Q - 0 1", 3),
]


@pytest.mark.parametrize("fen,depth", ROOTS)
def test_every_node_of_the_tree_matches_the_oracle(fen, depth):
    bad, visited, where = hostcheck.walk_compare(oracle.parse(fen), depth - 1)
    assert bad == 0, f"first mismatch at {where}"
    assert visited > 0


def test_random_playout_positions_match_the_oracle():
    """Differential check in the spirit of the reference's fuzz target
    (fuzz/fuzz_targets/generate_moves.rs:25-40): seeded random games."""
    total = 0
    start = oracle.starting()
    for g in range(300):
        arr = oracle.random_playout(start, 0x5EED0000 + g, 200)
        for i in range(len(arr)):
            p = oracle.Position.from_buffer_copy(arr[i].tobytes())
            bad, visited, where = hostcheck.walk_compare(p, 0)
            assert bad == 0, where
            total += visited
    assert total > 20000


def test_published_perft_values_through_the_device_logic():
    """tests/published_perft.py: independent values, through the leaf-kernel code path compiled for the host."""
    from published_perft import MAX_MOVES_FEN, PUBLISHED
    for fen, depth, nodes in PUBLISHED:
        if nodes <= 4_000_000:
            p = oracle.parse(fen)
            assert hostcheck.lib().hostcheck_perft(C.byref(p), depth) == nodes, fen
    p = oracle.parse(MAX_MOVES_FEN)
    assert hostcheck.lib().hostcheck_count_moves(C.byref(p)) == 218
    assert gen(p) == oracle.generate_moves(p)


def test_attack_info_matches_the_reference_pictures_and_the_oracle(vectors):
    """AttackInfo as the reference publishes it (attacks.rs:555-695 pictures), all five fields."""
    def device_ai(p):
        out = (C.c_uint64 * 5)()
        hostcheck.lib().hostcheck_attack_info(C.byref(p), out)
        return dict(zip(("attacks", "checkers", "pins", "xrays", "safe_king_squares"), out))
    for case in vectors["attacks"]["attack_info"]:
        got = device_ai(oracle.parse(case["fen"]))
        for field, bits in case["fields"].items():
            assert got[field] == int(bits), (case["source"], field)
    start = oracle.starting()
    n = 0
    for g in range(120):
        arr = oracle.random_playout(start, 0xA77AC0 + g, 160)
        for i in range(len(arr)):
            p = oracle.Position.from_buffer_copy(arr[i].tobytes())
            want = oracle.attack_info(p)
            got = device_ai(p)
            assert all(got[f] == getattr(want, f) for f in got), oracle.to_fen(p)
            n += 1
    assert n > 8000


def incremental(p, depth):
    out = (C.c_uint64 * 5)()
    hostcheck.lib().hostcheck_incremental(C.byref(p), depth, out)
    return tuple(out)


@pytest.mark.parametrize("fen,depth", ROOTS)
def test_quiet_move_fast_path_is_exact(fen, depth):
    """incremental.cuh: at every node of the tree taken as a K2 parent, (1) the clean/other split covers exactly
    the legal moves, (2) the counting sink agrees with the emitting one, (3) every clean move's fast reply count
    equals the full count of the child."""
    parents, moves, clean, mismatches, inert = incremental(oracle.parse(fen), max(0, depth - 2))
    assert mismatches == 0 and parents > 0 and moves >= clean >= inert


def test_quiet_move_fast_path_on_playout_positions():
    total = total_clean = 0
    for g in range(400):
        root = oracle.starting() if g % 2 == 0 else oracle.parse(__import__("workloads").chess324_fens()[g % 324])
        arr = oracle.random_playout(root, 0x1AC0000 + g, 200)
        for i in range(len(arr)):
            p = oracle.Position.from_buffer_copy(arr[i].tobytes())
            parents, moves, clean, mismatches, inert = incremental(p, 0)
            assert mismatches == 0, oracle.to_fen(p)
            total += moves
            total_clean += clean
    assert total > 500_000 and 0.05 < total_clean / total < 0.95


SPECIAL = [  # positions unlike anything in the reference trees or the playouts: many queens, bare castling rights, promotions
    "knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1",
    "kn1Q3q/pp5q/4Qq1q/Q3Q3/q2Q2Q1/1QQ2q1Q/q4qPP/qq1Q2NK w - - 0 1",
    "n1n5/PPPk4/8/8/8/8/4Kppp/5N1N b - - 0 1",
    "r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1",
    "4k3/8/8/8/8/8/8/4K2R w K - 0 1",
    "r3k2r/1b4bq/8/8/8/8/7B/R3K2R w KQkq - 0 1",
    "8/1n4N1/2k5/8/8/5K2/1N4n1/8 b - - 0 1",
    "rnbqkb1r/ppppp1pp/7n/4Pp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3",
    "8/8/1k6/2b5/2pP4/8/5K2/8 b - d3 0 1",
]


@pytest.mark.parametrize("fen", SPECIAL)
def test_quiet_move_fast_path_on_special_positions(fen):
    parents, moves, clean, mismatches, inert = incremental(oracle.parse(fen), 2)
    assert mismatches == 0 and parents > 50 and moves >= clean >= inert

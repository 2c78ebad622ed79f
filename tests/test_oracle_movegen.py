"""Oracle move generation, make_move, AttackInfo and perft against every vector the
reference's tests hold (tests/chess.rs:219-818, benches/chess.rs:52-86,
attacks.rs:555-695)."""
import pytest

import oracle

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"


def uci_moves(p):
    return [oracle.move_to_uci(m) for m in oracle.generate_moves(p)]


def test_move_list_vectors(vectors):
    assert len(vectors["move_lists"]) == 25
    for case in vectors["move_lists"]:
        assert sorted(uci_moves(oracle.parse(case["fen"]))) == case["moves"], case["source"]


def test_depth1_counts(vectors):
    for case in vectors["move_counts"]:
        assert len(uci_moves(oracle.parse(case["fen"]))) == case["count"], case["source"]


def test_move_order_follows_the_reference_code_order():
    # SURVEY.md Appendix A.3 (position.rs:331-406, LS1B-first iteration bitboard.rs:307-323)
    assert " ".join(uci_moves(oracle.parse(KIWIPETE))) == (
        "e1d1 e1f1 c3b1 c3d1 c3a4 c3b5 e5d3 e5c4 e5g4 e5c6 e5g6 e5d7 e5f7 a1b1 a1c1 a1d1 h1f1 h1g1 "
        "f3d3 f3e3 f3g3 f3h3 f3f4 f3f5 f3f6 d2c1 d2e3 d2f4 d2g5 d2h6 e2d1 e2f1 e2d3 e2c4 e2b5 e2a6 "
        "f3g4 f3h5 g2h3 d5e6 a2a3 b2b3 g2g3 d5d6 a2a4 g2g4 e1g1 e1c1")
    assert " ".join(uci_moves(oracle.starting())) == (
        "b1a3 b1c3 g1f3 g1h3 a2a3 b2b3 c2c3 d2d3 e2e3 f2f3 g2g3 h2h3 "
        "a2a4 b2b4 c2c4 d2d4 e2e4 f2f4 g2g4 h2h4")
    # promotions come out Q, R, B, N (position.rs:1133-1138)
    promo = uci_moves(oracle.parse("2n4k/1PP5/6K1/3Pp1Q1/3N4/3P4/P3R3/8 w - e6 0 1"))
    i = promo.index("b7c8q")
    assert promo[i:i + 4] == ["b7c8q", "b7c8r", "b7c8b", "b7c8n"]


def test_make_move_sequences(vectors):
    assert len(vectors["make_move"]) == 6
    for case in vectors["make_move"]:
        p = oracle.parse(case["start"])
        for step in case["steps"]:
            if "move" in step:
                oracle.make_move(p, oracle.move_from_uci(step["move"]))
            else:
                assert oracle.to_fen(p) == step["fen"], case["test"]


def test_hash_properties():
    """tests/chess.rs:820-868 (values are build-random in the reference; properties only)."""
    p = oracle.parse("8/5k2/6p1/8/8/8/1p3P2/5K2 w - - 0 1")
    h0 = p.hash
    for mv in ("f1e2", "f7f6", "e2f1"):
        oracle.make_move(p, oracle.move_from_uci(mv))
        assert p.hash != h0
    oracle.make_move(p, oracle.move_from_uci("f6f7"))
    assert oracle.to_fen(p) == "8/5k2/6p1/8/8/8/1p3P2/5K2 w - - 4 3" and p.hash == h0
    assert oracle.parse("6qk/8/8/3Pp3/8/8/K7/8 w - e6 0 1").hash != oracle.parse("6qk/8/8/3Pp3/8/8/K7/8 w - - 0 1").hash
    assert oracle.parse("6qk/8/8/3Pp3/8/8/K7/8 w - - 0 1").hash == oracle.parse("6qk/8/8/3Pp3/8/8/K7/8 w - - 2 3").hash


def test_conditional_en_passant_square():
    # position.rs:606-615: recorded only if an enemy pawn attacks the skipped square.
    p = oracle.starting()
    oracle.make_move(p, oracle.move_from_uci("e2e4"))
    assert p.en_passant_square == 64
    p = oracle.parse("rnbqkbnr/ppp1pppp/8/8/3p4/8/PPPPPPPP/RNBQKBNR w KQkq - 0 3")
    oracle.make_move(p, oracle.move_from_uci("e2e4"))
    assert oracle.to_fen(p).split()[3] == "e3"


def test_attack_info_pictures(vectors):
    assert len(vectors["attacks"]["attack_info"]) == 4
    for case in vectors["attacks"]["attack_info"]:
        ai = oracle.attack_info(oracle.parse(case["fen"]))
        for field, bits in case["fields"].items():
            assert getattr(ai, field) == int(bits), (case["source"], field)


def test_in_check():
    assert oracle.in_check(oracle.parse("3kn3/3p4/8/6B1/8/6K1/3R4/8 b - - 0 1"))
    assert not oracle.in_check(oracle.starting())


def test_perft_kats(vectors):
    """Every perft assertion in tests/chess.rs (fast and #[ignore]d) and benches/chess.rs."""
    cases = [(c["fen"], c["depth"], c["nodes"], c["source"]) for c in vectors["perft"] + vectors["bench_perft"]]
    assert len(cases) == 78
    checked = 0
    for fen, depth, nodes, source in cases:
        threads = 8 if nodes > 50_000_000 else 1
        if nodes > 2_000_000_000:
            continue  # CPW challenge depth 7: test_perft_cpw_challenge (slow)
        assert oracle.perft(oracle.parse(fen), depth, threads) == nodes, source
        checked += 1
    assert checked == 77


@pytest.mark.slow
def test_perft_cpw_challenge(vectors):
    case = [c for c in vectors["perft"] if c["nodes"] > 2_000_000_000]
    assert len(case) == 1 and case[0]["nodes"] == 14_794_751_816      # tests/chess.rs:815-818
    assert oracle.perft(oracle.parse(case[0]["fen"]), case[0]["depth"], 8) == case[0]["nodes"]


def test_perft_mt_equals_single_thread():
    p = oracle.parse(KIWIPETE)
    assert oracle.perft(p, 4, 1) == oracle.perft(p, 4, 4) == 4_085_603


def test_batch_and_playout_helpers():
    start = oracle.starting()
    a = oracle.random_playout(start, 0x5EED0000, 60)
    b = oracle.random_playout(start, 0x5EED0000, 60)
    assert len(a) > 10 and a.tobytes() == b.tobytes()
    offsets, moves = oracle.generate_moves_batch(a)
    assert offsets[0] == 0 and offsets[-1] == len(moves)
    for i in (0, 1, len(a) - 1):
        p = oracle.Position.from_buffer_copy(a[i].tobytes())
        assert list(moves[offsets[i]:offsets[i + 1]]) == oracle.generate_moves(p)


def test_workload_fixtures_are_the_oracles_answers():
    """tests/golden/workload_fixtures.json (what bench.py checks the GPU's other_configs against) is reproducible from the
    oracle: same digests for the stand-in book and the first 100 000 playout positions' worth of the 1 M workload."""
    import hashlib
    import json
    from pathlib import Path
    import numpy as np
    import workloads
    fx = json.loads((Path(__file__).resolve().parent / "golden" / "workload_fixtures.json").read_text())
    book = workloads.chess324_synth_book()
    assert hashlib.sha256(book.tobytes()).hexdigest() == fx["book"]["sha256"] and len(fx["book"]["perft"]) == len(book) == 1620
    for i in range(0, len(book), 97):
        assert oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), 4) == fx["book"]["perft"][i]
    assert sum(fx["book"]["perft"]) == fx["book"]["nodes"]
    positions = workloads.playout_positions(1_000_000)
    assert hashlib.sha256(positions.tobytes()).hexdigest() == fx["playouts"]["sha256"]
    offsets, moves = oracle.generate_moves_batch(positions)
    assert len(moves) == fx["playouts"]["moves"]
    assert hashlib.sha256(offsets.tobytes()).hexdigest() == fx["playouts"]["offsets_sha256"]
    assert hashlib.sha256(moves.tobytes()).hexdigest() == fx["playouts"]["moves_sha256"]

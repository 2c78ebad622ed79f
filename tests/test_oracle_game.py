"""The oracle's `Game` against the reference's own game tests (src/chess/game.rs:143-227; the
tablebase tests need shakmaty-syzygy and are out of scope)."""
import oracle


def play(game, *ucis):
    for u in ucis:
        game.apply(oracle.move_from_uci(u))


def test_detect_repetition():
    """game.rs:143-167"""
    g = oracle.Game(oracle.starting())
    assert g.result() is None
    for i, u in enumerate(["g1f3", "g8f6", "f3g1", "f6g8", "g1f3", "g8f6", "f3g1"]):
        play(g, u)
        assert g.result() is None, i
    play(g, "f6g8")
    assert g.result() == "draw"


def test_stalemate():
    """game.rs:189-200"""
    g = oracle.Game(oracle.from_fen("3b2qk/p6p/1p3Q1P/8/8/n7/PP6/K7 b - - 3 2"))
    assert g.result() is None
    play(g, "d8f6")
    assert g.actions() == [] and g.result() == "draw"


def test_checkmate():
    """game.rs:202-213: the root's player (white) delivers mate -> Win"""
    g = oracle.Game(oracle.from_fen("3b3k/p5qp/1p3Q1P/8/8/n7/PP6/K7 w - - 4 3"))
    assert g.result() is None
    play(g, "f6g7")
    assert g.actions() == [] and g.result() == "win"


def test_fifty_move_rule():
    """game.rs:215-227"""
    g = oracle.Game(oracle.from_fen("8/5k2/3p4/1p1Pp2p/pP2Pp1P/P4P1K/8/8 b - - 99 50"))
    assert g.result() is None
    play(g, "f7f6")
    assert g.result() == "draw"


def test_fourth_occurrence_is_not_flagged():
    """RepetitionTable::record returns true exactly at the third occurrence (zobrist.rs:28-32)."""
    g = oracle.Game(oracle.starting())
    cycle = ["g1f3", "g8f6", "f3g1", "f6g8"]
    play(g, *cycle, *cycle)
    assert g.result() == "draw"
    for u in cycle[:3]:        # third occurrences of the three intermediate positions
        play(g, u)
        assert g.result() == "draw"
    play(g, cycle[3])          # fourth occurrence of the start position: count == 4, not 3
    assert g.result() is None


def test_random_games_terminate_with_a_result():
    seen = set()
    for s in range(60):
        g = oracle.Game(oracle.starting())
        r, plies = g.play_random(0x6A3E0000 + s, 600)
        assert r in ("win", "draw", "loss", None) and plies <= 600
        assert (r is None) == (plies == 600 and g.result() is None)
        seen.add(r)
    assert "draw" in seen


TRIANGULATION = ["d3c3", "d5e5", "c3c2", "e5d5", "c2d3",          # same placement as the root, black to move
                 "d5e5", "d3c3", "e5d5", "c3c2", "d5e5", "c2d3", "e5d5"]


def test_repetition_key_ignores_the_side_to_move_like_the_reference():
    """Position::make_move never toggles the side-to-move key and never removes an en passant key
    (position.rs:414-433, :606-615; only compute_hash uses BLACK_TO_MOVE, :780), so the reference's
    RepetitionTable counts equal placement + castling rights.  After a king triangulation the root
    placement occurs a third time at ply 12 although the side to move matches the root only twice."""
    g = oracle.Game(oracle.from_fen("8/8/8/3k4/8/3K4/8/8 w - - 0 1"))
    for i, u in enumerate(TRIANGULATION[:-1]):
        play(g, u)
        assert g.result() is None, i
    play(g, TRIANGULATION[-1])
    assert g.result() == "draw"

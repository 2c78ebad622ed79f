"""Oracle FEN/EPD parsing, printing and validation against tests/chess.rs:33-181."""
import pytest

import oracle


def sanitize_fen(text: str) -> str:
    """tests/chess.rs:12-26"""
    text = text.strip()
    for prefix in ("fen ", "epd "):
        if text.startswith(prefix):
            text = text[len(prefix):]
    return text if len(text.split()) == 6 else text + " 0 1"


def test_starting(vectors):
    assert oracle.to_fen(oracle.starting()) == vectors["starting_fen"]


def test_basic_positions_round_trip(vectors):
    assert len(vectors["fen"]["roundtrip"]) == 9
    for fen in vectors["fen"]["roundtrip"]:
        assert oracle.to_fen(oracle.from_fen(fen)) == sanitize_fen(fen)


def test_validation_error_strings(vectors):
    assert len(vectors["fen"]["errors"]) == 13
    for case in vectors["fen"]["errors"]:
        with pytest.raises(oracle.OracleError) as e:
            oracle.parse(case["fen"])
        assert case["expected"] in str(e.value), case
        assert str(e.value).startswith("illegal position")   # position.rs:273


def test_clean_board_str_and_no_crash(vectors):
    for text in vectors["fen"]["try_from_ok"]:
        oracle.parse(text)
    for text in vectors["fen"]["try_from_err"]:
        with pytest.raises(oracle.OracleError):
            oracle.parse(text)
    for text in vectors["fen"]["from_fen_err"]:
        with pytest.raises(oracle.OracleError):
            oracle.from_fen(text)


@pytest.mark.parametrize("fen", [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0",        # halfmove without fullmove
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 0",      # fullmove 0
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 256 1",    # halfmove > u8
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1 x",    # trailing
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w QKkq - 0 1",      # castling order
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR  w KQkq - 0 1",     # double space
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP w KQkq - 0 1",               # 7 ranks
    "rnbqkbnr/pppppppp/8/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",    # 9 ranks
    "rnbqkbnr/pppppppp/44/8/8/8/PPPPPPP/RNBQKBNR w KQkq - 0 1",      # short rank
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq e9 0 1",     # bad ep
    "", " ", "w", "8/8/8/8/8/8/8/8",
])
def test_appendix_c_rejections(fen):
    with pytest.raises(oracle.OracleError):
        oracle.from_fen(fen)


def test_fen_supplied_ep_square_kept_verbatim():
    # tests/chess.rs:45 -- no black pawn can capture on c3, the square still round-trips.
    fen = "8/8/8/8/2P5/3k4/8/KB6 b - c3 0 1"
    assert oracle.to_fen(oracle.from_fen(fen)) == fen


def test_move_uci_round_trip():
    for uci in ("e2e4", "a7a8q", "h2h1n", "b7c8r", "e1g1", "g7g8b"):
        assert oracle.move_to_uci(oracle.move_from_uci(uci)) == uci
    # core.rs:24-66 packing
    assert oracle.move_from_uci("e2e4") == 12 | (28 << 6)
    assert oracle.move_from_uci("a7a8q") == 48 | (56 << 6) | (4 << 12)
    assert oracle.move_from_uci("a7a8n") >> 12 == 1

"""CPU tests of the shipped library's host side: the C ABI loads and exports every
symbol include/pabi_cuda.h declares, and the text API (Position::starting /
from_fen / TryFrom / Display, Move text) reproduces tests/chess.rs:33-181 and
matches the oracle.  No compute entry point is called here (there is no GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import oracle
import pabi_b200 as pb
from pabi_b200 import _native

ROOT = Path(__file__).resolve().parent.parent


def sanitize_fen(text: str) -> str:
    """tests/chess.rs:12-26"""
    text = text.strip()
    for prefix in ("fen ", "epd "):
        if text.startswith(prefix):
            text = text[len(prefix):]
    return text if len(text.split()) == 6 else text + " 0 1"


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "pabi_cuda.h").read_text()
    declared = set(re.findall(r"\b(pabi_(?:cuda|position|move)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    L = _native.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert C.sizeof(_native.PositionStruct) == 112 and C.sizeof(_native.Timing) == 80


def test_rust_binding_declares_every_entry_point():
    """rust/ffi.rs (the `extern "C"` block INTEGRATION.md asks a pabi maintainer to add; not compiled here, there is no
    rustc) names exactly the functions of include/pabi_cuda.h, with the same number of parameters."""
    def signatures(text, pattern):
        out = {}
        for name, args in re.findall(pattern, text, flags=re.S):
            args = re.sub(r"/\*.*?\*/", "", args, flags=re.S).strip()
            out[name] = 0 if args in ("", "void") else args.count(",") + 1
        return out
    header = signatures((ROOT / "include" / "pabi_cuda.h").read_text(), r"\b(pabi_(?:cuda|position|move)_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;")
    rust = signatures((ROOT / "rust" / "ffi.rs").read_text(), r"\bfn\s+(pabi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*(?:->[^;{]*)?;")
    assert set(header) == set(rust)
    assert header == rust


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pb.PabiCudaError):
        pb.Engine()
    with pytest.raises(pb.PabiCudaError):
        pb.perft(pb.Position.starting(), 1)


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "hostcheck_" not in out


def test_starting(vectors):
    assert str(pb.Position.starting()) == vectors["starting_fen"]
    assert bytes(pb.Position.starting().raw)[:96] == bytes(oracle.starting())[:96]


def test_basic_positions_round_trip(vectors):
    for fen in vectors["fen"]["roundtrip"]:
        p = pb.Position.from_fen(fen)
        assert str(p) == sanitize_fen(fen)
        o = oracle.from_fen(fen)
        o.hash = 0
        assert bytes(p.raw) == bytes(o)


def test_validation_error_strings(vectors):
    for case in vectors["fen"]["errors"]:
        with pytest.raises(ValueError) as e:
            pb.Position.try_from(case["fen"])
        assert case["expected"] in str(e.value), case
        assert str(e.value).startswith("illegal position")
        with pytest.raises(oracle.OracleError) as o:
            oracle.parse(case["fen"])
        assert str(o.value) == str(e.value)


def test_clean_board_str_and_no_crash(vectors):
    for text in vectors["fen"]["try_from_ok"]:
        assert str(pb.Position.try_from(text)) == oracle.to_fen(oracle.parse(text))
    for text in vectors["fen"]["try_from_err"]:
        with pytest.raises(ValueError):
            pb.Position.try_from(text)
    for text in vectors["fen"]["from_fen_err"]:
        with pytest.raises(ValueError):
            pb.Position.from_fen(text)


REJECTED = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 0",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 256 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1 x",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w QKkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KK - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w  - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR  w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/44/8/8/8/PPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/9/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8p/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq e9 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR x KQkq - 0 1",
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNX w KQkq - 0 1",
    "", " ", "w", "8/8/8/8/8/8/8/8",
]


@pytest.mark.parametrize("fen", REJECTED)
def test_rejections_with_the_oracles_message(fen):
    with pytest.raises(ValueError) as e:
        pb.Position.from_fen(fen)
    with pytest.raises(oracle.OracleError) as o:
        oracle.from_fen(fen)
    assert str(e.value) == str(o.value)


def test_accepts_what_the_oracle_accepts_on_playout_positions():
    start = oracle.starting()
    for g in range(40):
        arr = oracle.random_playout(start, 0xABC0 + g, 150, 3)
        for i in range(len(arr)):
            o = oracle.Position.from_buffer_copy(arr[i].tobytes())
            fen = oracle.to_fen(o)
            p = pb.Position.from_fen(fen)
            assert str(p) == fen
            o.hash = 0
            assert bytes(p.raw) == bytes(o)


def test_fen_supplied_ep_square_kept_verbatim():
    fen = "8/8/8/8/2P5/3k4/8/KB6 b - c3 0 1"  # tests/chess.rs:45
    assert str(pb.Position.from_fen(fen)) == fen


def test_move_text():
    for uci in ("e2e4", "a7a8q", "h2h1n", "b7c8r", "e1g1", "g7g8b"):
        m = pb.Move.from_uci(uci)
        assert str(m) == uci and m.raw == oracle.move_from_uci(uci)
    assert pb.Move.from_uci("e2e4").raw == 12 | (28 << 6)     # core.rs:24-66
    assert pb.Move(48, 56, "q").raw == 48 | (56 << 6) | (4 << 12)
    m = pb.Move.from_uci("a7a8n")
    assert (m.from_(), m.to(), m.promotion()) == (48, 56, "n")
    for bad in ("", "e2", "e2e9", "e2e4x", "e2e4qq", "i2e4"):
        with pytest.raises(ValueError):
            pb.Move.from_uci(bad)


def test_pack64_round_trip():
    """The 64-byte record as a wire format: pack / unpack reproduces the position (hash is not carried)."""
    for g in range(10):
        arr = oracle.random_playout(oracle.starting(), 0x64640000 + g, 120, 5)
        for i in range(len(arr)):
            o = oracle.Position.from_buffer_copy(arr[i].tobytes())
            p = pb.Position.from_fen(oracle.to_fen(o))
            words = p.pack64()
            assert len(words) == 8
            q = pb.Position.unpack64(words)
            assert bytes(q.raw) == bytes(p.raw)

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
    declared = set(re.findall(r"\b(pabi_(?:cuda|position|move|zobrist)_[a-z0-9_]+)\s*\(", header))
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
    header = signatures((ROOT / "include" / "pabi_cuda.h").read_text(), r"\b(pabi_(?:cuda|position|move|zobrist)_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;")
    rust = signatures((ROOT / "rust" / "ffi.rs").read_text(), r"\bfn\s+(pabi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*(?:->[^;{/]*)?;")
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
        assert bytes(p.raw) == bytes(oracle.from_fen(fen))  # all 112 bytes: Position::hash is compute_hash on both sides


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
            assert bytes(p.raw) == bytes(oracle.from_fen(fen))  # the parser caches compute_hash (position.rs:266)


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


# ---- Position::hash and the small accessors of the boundary ----

def _splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def test_zobrist_keys_follow_the_documented_recipe():
    """include/pabi_cuda.h: 776 splitmix64 outputs from the seed "pabi", then the reference's five constants
    (src/chess/generated.rs:7-13)."""
    keys = pb.chess.zobrist_keys()
    state = 0x70616269
    for i in range(776):
        state, want = _splitmix64(state)
        assert int(keys[i]) == want, i
    assert [int(k) for k in keys[776:]] == [0x9E06BAD39D761293, 0xF05AC573DD61D323, 0x41D8B55BA5FEB78B,
                                            0x680988787A43D289, 0x2F941F8DFD3E3D1F]
    assert len(set(int(k) for k in keys)) == 781


def test_hash_of_parsed_positions_is_compute_hash_and_equals_the_oracles(vectors):
    import hostcheck
    fens = list(vectors["fen"]["roundtrip"]) + [c["fen"] for c in vectors["perft"] if c["fen"] != "startpos"]
    for fen in fens:
        p, o = pb.Position.try_from(fen), oracle.parse(fen)
        assert p.hash() == p.compute_hash() == o.hash == oracle.lib().oracle_compute_hash(C.byref(o))
        assert hostcheck.lib().hostcheck_compute_hash(C.byref(o)) == o.hash  # the device function, compiled for the host


def test_hash_properties_of_the_reference():
    """tests/chess.rs:840-868: the parts that need no make_move (en_passant_hash, castling_hash's first assertion,
    move_clock_hash); the make_move parts run on the device (tests/test_gpu_parity.py)."""
    h = lambda fen: pb.Position.try_from(fen).hash()
    assert h("6qk/8/8/3Pp3/8/8/K7/8 w - e6 0 1") != h("6qk/8/8/3Pp3/8/8/K7/8 w - - 0 1")
    assert h("rnbqk1nr/p3bppp/1p2p3/2ppP3/3P4/P7/1PP1NPPP/R1BQKBNR w KQkq - 0 7") != \
        h("rnbqk1nr/p3bppp/1p2p3/2ppP3/3P4/P7/1PP1NPPP/R1BQKBNR w Qkq - 0 7")
    assert h("6qk/8/8/3Pp3/8/8/K7/8 w - - 0 1") == h("6qk/8/8/3Pp3/8/8/K7/8 w - - 2 3")
    assert h("6qk/8/8/3Pp3/8/8/K7/8 w - - 0 1") != h("6qk/8/8/3Pp3/8/8/K7/8 b - - 0 1")


def test_small_accessors():
    start = pb.Position.starting()
    assert start.num_pieces() == 32 and not start.halfmove_clock_expired()                  # position.rs:122-124
    late = pb.Position.from_fen("8/5k2/3p4/1p1Pp2p/pP2Pp1P/P4P1K/8/8 b - - 100 80")
    assert late.num_pieces() == 14 and late.halfmove_clock_expired()                       # position.rs:750-752
    assert not pb.Position.from_fen("8/5k2/3p4/1p1Pp2p/pP2Pp1P/P4P1K/8/8 b - - 99 80").halfmove_clock_expired()
    # core.rs:83-90 doc-test: E1 -> E8 flipped is E8 -> E1; the promotion is kept
    assert pb.Move(4, 60).flip_perspective() == pb.Move(60, 4)
    assert pb.Move.from_uci("a7a8q").flip_perspective() == pb.Move.from_uci("a2a1q")
    assert pb.Move.from_uci("d4d5").flip_perspective() == pb.Move.from_uci("d5d4")          # core.rs:227-228


def test_c_consumer_of_the_header(tmp_path):
    """tests/native/abi_smoke.c: a C11 translation unit that includes include/pabi_cuda.h, asserts both struct layouts
    with offsetof and links the library (host entry points only here; with a device it also runs one perft)."""
    import subprocess
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests/native/abi_smoke.c"),
                    "-o", str(exe), "-L", str(_native.PKG_DIR), "-lpabi_b200", f"-Wl,-rpath,{_native.PKG_DIR}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "layout ok" in out.stdout and "text api ok" in out.stdout

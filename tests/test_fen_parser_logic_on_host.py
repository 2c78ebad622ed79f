"""The device FEN / EPD parser (pabi_b200/csrc/fen.cuh, compiled for the host) against the library's
host parser (= the reference's rules, tests/chess.rs:33-181) on the reference's vectors, on playout
positions, and on randomly mutated lines: same accept / reject decision, same position."""
import ctypes as C
import random

import hostcheck
import oracle
import pabi_b200 as pb
from test_host_api import REJECTED


def device_parse(text: str):
    raw = text.encode("utf-8")
    p = oracle.Position()
    rc = hostcheck.lib().hostcheck_parse(raw, len(raw), C.byref(p))
    return rc, p


def agree(text: str):
    rc, p = device_parse(text)
    try:
        want = pb.Position.try_from(text)
    except ValueError as e:
        assert rc < 0, (text, str(e))
        return rc
    assert rc == 0, (text, rc)
    assert bytes(p) == bytes(want.raw), text
    return 0


def test_reference_vectors(vectors):
    fen = vectors["fen"]
    for text in fen["roundtrip"] + fen["try_from_ok"]:
        assert agree(text) == 0
    for text in fen["try_from_err"] + REJECTED:
        assert agree(text) < 0
    for case in fen["errors"]:
        assert agree(case["fen"]) == -8           # validate: "illegal position"
    assert agree("\n epd rnbqkb1r/ppp1pp1p/5np1/3p4/3P1B2/5N2/PPP1PPPP/RN1QKB1R w KQkq -\n") == 0   # try_from trims


def test_status_classes():
    base = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"
    for text, rc in ((base + "x w KQkq - 0 1", -1), (base + " x KQkq - 0 1", -2), (base + " w QK - 0 1", -3),
                     (base + " w KQkq e9 0 1", -4), (base + " w KQkq - 256 1", -5), (base + " w KQkq - 0 0", -6),
                     (base + " w KQkq - 0", -6), (base + " w KQkq - 0 1 x", -7), ("3k4/8/8/8/8/8/8/8 w - - 0 1", -8)):
        assert agree(text) == rc, text


def test_playout_positions_and_mutations():
    rng = random.Random(11)
    alphabet = "pnbrqkPNBRQK12345678/ wbKQkq-abcdefgh09+"
    checked = accepted = 0
    for g in range(60):
        arr = oracle.random_playout(oracle.starting(), 0xFE110000 + g, 160, 2)
        for i in range(len(arr)):
            fen = oracle.to_fen(oracle.Position.from_buffer_copy(arr[i].tobytes()))
            assert agree(fen) == 0
            assert agree(" ".join(fen.split()[:4])) == 0          # EPD without operations
            for _ in range(6):                                    # random edits: same verdict as the host parser
                chars = list(fen)
                k = rng.randrange(len(chars))
                op = rng.randrange(3)
                if op == 0:
                    chars[k] = rng.choice(alphabet)
                elif op == 1:
                    del chars[k]
                else:
                    chars.insert(k, rng.choice(alphabet))
                accepted += agree("".join(chars)) == 0
                checked += 1
    assert checked > 20000 and 0 < accepted < checked

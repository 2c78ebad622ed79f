#!/usr/bin/env python3
"""Extracts the reference's own golden vectors for the perft / move-generation path.

Run in the build container, where the reference is mounted read-only:

    python tests/golden/make_golden.py [/root/reference]

Outputs (committed; the GPU box has no /root/reference):

* tests/golden/reference_tables.npz   -- every table present under
  /root/reference/generated/*.rs, as uint64 arrays (generated.rs:30-82).
  generated/rook_attacks.rs is absent from the mount (.MISSING_LARGE_BLOBS).
* tests/golden/reference_vectors.json -- the known-answer data of
  /root/reference/tests/chess.rs, benches/chess.rs, src/chess/attacks.rs (tests),
  src/chess/game.rs:179-181 and the doc-tests, each with its source line.

Only data (FEN strings, move strings, node counts, bitboard pictures, error
substrings) is extracted; no reference code is copied.
"""
import json
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent


def line_of(text: str, pos: int) -> int:
    return text.count("\n", 0, pos) + 1


def unescape(s: str) -> str:
    """Rust string literal body -> text (handles `\\` line continuations)."""
    s = re.sub(r"\\\n\s*", "", s)
    return s.replace('\\"', '"').replace("\\n", "\n").replace("\\t", "\t")


STR = r'"((?:[^"\\]|\\.)*)"'


def tables() -> dict:
    out = {}
    for path in sorted((REF / "generated").glob("*.rs")):
        body = path.read_text()
        vals = [int(v.replace("_", ""), 16) for v in re.findall(r"0x([0-9a-fA-F_]+)", body)]
        if not vals:  # offsets / relevant-occupancy files hold decimal values (with `_` separators)
            vals = [int(v.replace("_", "")) for v in re.findall(r"\b(\d[\d_]*)\b", body)]
        out[path.stem] = np.array(vals, dtype=np.uint64)
    return out


def perft_vectors(text: str, file: str) -> list:
    """`let position = setup("FEN") | Position::starting();` followed by
    `assert_eq!(perft(&position, D), N);` lines, per #[test] function."""
    out = []
    for fn in re.finditer(r"((?:#\[[^\]]*\]\s*)+)fn (\w+)\(\) \{(.*?)\n\}", text, re.S):
        attrs, name, body = fn.group(1), fn.group(2), fn.group(3)
        fen = None
        m = re.search(r"let (?:mut )?position\s*=\s*(?:setup\(\s*" + STR + r"\s*,?\s*\)|Position::starting\(\))", body, re.S)
        if not m:
            continue
        fen = unescape(m.group(1)) if m.group(1) is not None else "startpos"
        for a in re.finditer(r"assert_eq!\(perft\(&position, (\d+)\), ([\d_]+)\);", body):
            out.append({
                "fen": fen, "depth": int(a.group(1)), "nodes": int(a.group(2).replace("_", "")),
                "ignored": "#[ignore]" in attrs, "test": name,
                "source": f"{file}:{line_of(text, fn.start(3) + a.start())}",
            })
    return out


def bench_vectors(text: str) -> list:
    out = []
    start = text.index("let test_cases = vec![")
    end = text.index("];", start)
    body = text[start:end]
    for m in re.finditer(r"\(\s*(Position::starting\(\)|Position::from_fen\(\s*" + STR + r"\s*,?\s*\)\s*\.unwrap\(\))\s*,\s*(\d+)\s*,\s*([\d_]+)\s*,?\s*\)", body, re.S):
        fen = unescape(m.group(2)) if m.group(2) is not None else "startpos"
        out.append({"fen": fen, "depth": int(m.group(3)), "nodes": int(m.group(4).replace("_", "")),
                    "source": f"benches/chess.rs:{line_of(text, start + m.start())}"})
    return out


def movelist_vectors(text: str) -> list:
    out = []
    pat = (r"get_moves\(&(?:setup\(\s*" + STR + r"\s*,?\s*\)|Position::starting\(\))\)\s*,\s*"
           r"sorted_moves\(&\[(.*?)\]\)")
    for m in re.finditer(pat, text, re.S):
        fen = unescape(m.group(1)) if m.group(1) is not None else "startpos"
        moves = re.findall(r'"([a-h][1-8][a-h][1-8][nbrq]?)"', m.group(2))
        out.append({"fen": fen, "moves": sorted(moves), "source": f"tests/chess.rs:{line_of(text, m.start())}"})
    return out


def movecount_vectors(text: str) -> list:
    out = []
    pat = r"get_moves\(&setup\(\s*" + STR + r"\s*,?\s*\)\)\s*\.len\(\)\s*,\s*(\d+)"
    for m in re.finditer(pat, text, re.S):
        out.append({"fen": unescape(m.group(1)), "count": int(m.group(2)),
                    "source": f"tests/chess.rs:{line_of(text, m.start())}"})
    return out


def make_move_vectors(text: str) -> list:
    """Per test: start position, then interleaved make_move / to_string assertions."""
    out = []
    for name in ("make_moves", "basic_moves", "promotion_moves", "castling_reset", "repetition_hash",
                 "castling_hash"):
        fn = re.search(r"fn " + name + r"\(\) \{(.*?)\n\}", text, re.S)
        body = fn.group(1)
        m = re.search(r"let mut position = (?:setup\(\s*" + STR + r"\s*,?\s*\)|Position::starting\(\))", body, re.S)
        fen = unescape(m.group(1)) if m.group(1) is not None else "startpos"
        steps = []
        tok = re.compile(r'Move::from_uci\("(\w+)"\)|assert_eq!\(\s*position\.to_string\(\),\s*' + STR, re.S)
        for t in tok.finditer(body, m.end()):
            if t.group(1):
                steps.append({"move": t.group(1)})
            else:
                steps.append({"fen": unescape(t.group(2))})
        out.append({"test": name, "start": fen, "steps": steps,
                    "source": f"tests/chess.rs:{line_of(text, fn.start())}"})
    return out


def fen_vectors(text: str) -> dict:
    rt = [unescape(m.group(1)) for m in re.finditer(r"expect_legal_position\(\s*" + STR, text, re.S)]
    errors = []
    for m in re.finditer(r"#\[should_panic\(\s*expected = " + STR + r"\s*\)\]\s*fn (\w+)\(\) \{(.*?)\n\}", text, re.S):
        fen = re.search(r"(?:setup|Position::try_from)\(" + STR + r"\)", m.group(3), re.S)
        errors.append({"fen": unescape(fen.group(1)), "expected": unescape(m.group(1)), "test": m.group(2),
                       "source": f"tests/chess.rs:{line_of(text, m.start())}"})
    accept, reject, reject_from_fen = [], [], []
    for m in re.finditer(r"assert!\(\s*Position::(try_from|from_fen)\(\s*" + STR + r"\s*,?\s*\)\s*\.(is_ok|is_err)\(\)", text, re.S):
        s = unescape(m.group(2))
        if m.group(1) == "from_fen":
            (reject_from_fen if m.group(3) == "is_err" else accept).append(s)
        else:
            (accept if m.group(3) == "is_ok" else reject).append(s)
    return {"roundtrip": rt, "errors": errors, "try_from_ok": accept, "try_from_err": reject,
            "from_fen_err": reject_from_fen}


def picture_to_bits(picture: str) -> int:
    rows = [r.split() for r in picture.strip().split("\n")]
    assert len(rows) == 8 and all(len(r) == 8 for r in rows), picture
    bits = 0
    for i, row in enumerate(rows):  # first row is rank 8 (bitboard.rs:177-206)
        rank = 7 - i
        for file, c in enumerate(row):
            if c == "1":
                bits |= 1 << (rank * 8 + file)
    return bits


def attack_vectors(text: str) -> dict:
    out = {"attack_info": [], "tables": [], "rays": [], "empty_rays": []}
    # AttackInfo pictures (attacks.rs:555-695)
    for fn in re.finditer(r"fn (basic_attack_info|xrays|rich_attack_info|complicated_attack_info)\(\) \{(.*?)\n    \}", text, re.S):
        body = fn.group(2)
        fen = unescape(re.search(r"Position::try_from\(\s*" + STR, body, re.S).group(1))
        entry = {"fen": fen, "fields": {}, "source": f"src/chess/attacks.rs:{line_of(text, fn.start())}"}
        for a in re.finditer(r'format!\("\{:\?\}", attacks\.(\w+)\),\s*' + STR, body, re.S):
            entry["fields"][a.group(1)] = str(picture_to_bits(unescape(a.group(2))))
        for a in re.finditer(r"assert!\(attacks\.(\w+)\.is_empty\(\)\)", body):
            entry["fields"][a.group(1)] = "0"
        out["attack_info"].append(entry)
    # leaper pictures (attacks.rs:304-484)
    for a in re.finditer(r'format!\("\{:\?\}", (king|knight)_attacks\(Square::(\w\d)\)\),\s*' + STR, text, re.S):
        out["tables"].append({"table": a.group(1) + "_attacks", "square": a.group(2).lower(),
                              "bits": str(picture_to_bits(unescape(a.group(3))))})
    for a in re.finditer(r'format!\("\{:\?\}", pawn_attacks\(Square::(\w\d), Player::(White|Black)\)\),\s*' + STR, text, re.S):
        out["tables"].append({"table": a.group(2).lower() + "_pawn_attacks", "square": a.group(1).lower(),
                              "bits": str(picture_to_bits(unescape(a.group(3))))})
    for a in re.finditer(r'format!\("\{:\?\}", ray\(Square::(\w\d), Square::(\w\d)\)\),\s*' + STR, text, re.S):
        out["rays"].append({"from": a.group(1).lower(), "to": a.group(2).lower(),
                            "bits": str(picture_to_bits(unescape(a.group(3))))})
    for a in re.finditer(r"assert!\(ray\(Square::(\w\d), Square::(\w\d)\)\.is_empty\(\)\)", text):
        out["empty_rays"].append({"from": a.group(1).lower(), "to": a.group(2).lower()})
    # sliders KAT (attacks.rs:224-302)
    fn = re.search(r"fn sliders\(\) \{(.*?)\n    \}", text, re.S).group(1)
    occ_squares = re.findall(r"Square::(\w\d)", fn[: fn.index("]);")])
    pics = [unescape(p) for p in re.findall(r'assert_eq!\(\s*format!\(.*?\),\s*' + STR, fn, re.S)]
    out["sliders"] = {
        "occupancy_squares": [s.lower() for s in occ_squares],
        "occupancy": str(picture_to_bits(pics[0])),
        "bishop_relevant_e4": str(picture_to_bits(pics[1])),
        "bishop_attacks_e4": str(picture_to_bits(pics[2])),
        "rook_relevant_e4": str(picture_to_bits(pics[3])),
        "rook_attacks_e4": str(picture_to_bits(pics[4])),
        "source": "src/chess/attacks.rs:224-302",
    }
    return out


def main() -> None:
    np.savez_compressed(OUT / "reference_tables.npz", **tables())
    chess = (REF / "tests/chess.rs").read_text()
    bench = (REF / "benches/chess.rs").read_text()
    attacks = (REF / "src/chess/attacks.rs").read_text()
    vectors = {
        "generated_from": "kirillbobyrev/pabi v2025.3.13 (tests/chess.rs, benches/chess.rs, src/chess/attacks.rs)",
        "perft": perft_vectors(chess, "tests/chess.rs"),
        "bench_perft": bench_vectors(bench),
        "move_lists": movelist_vectors(chess),
        "move_counts": movecount_vectors(chess),
        "make_move": make_move_vectors(chess),
        "fen": fen_vectors(chess),
        "attacks": attack_vectors(attacks),
        # position.rs:68-76 doc-test, core.rs:149-157 doc-test
        "starting_fen": "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
        "square_numbering": {"a1": 0, "e1": 4, "h1": 7, "a4": 24, "h8": 63},
    }
    (OUT / "reference_vectors.json").write_text(json.dumps(vectors, indent=1) + "\n")
    for key in ("perft", "bench_perft", "move_lists", "move_counts", "make_move"):
        print(key, len(vectors[key]))
    print({k: len(v) for k, v in vectors["fen"].items()})
    print({k: (len(v) if isinstance(v, list) else 1) for k, v in vectors["attacks"].items()})


if __name__ == "__main__":
    main()

"""Writes tests/golden/workload_fixtures.json: the CPU oracle's answers for the seeded stand-in workloads of
SURVEY.md 8(d), so that `bench.py` can check what the GPU produced for BASELINE.json configs[2] and configs[4]
without executing the oracle (which it may only do in its cpu_baseline leg).

    python tests/golden/make_workload_fixtures.py

  book      'chess324-synth': sha256 of the 1 620 positions (112-byte records, Position::hash included) and the
            oracle's perft(4) of every one of them
  playouts  the first 1 000 000 positions of the seeded playouts: sha256 of the records, of the CSR offsets
            (u32) and of the moves (u16) of the oracle's ordered move lists, and the number of moves
tests/test_gpu_parity.py checks that the device-side generators (pabi_b200/workloads.py) reproduce both.
"""
import hashlib
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle  # noqa: E402
import workloads  # noqa: E402

book = workloads.chess324_synth_book()
counts = [oracle.perft(oracle.Position.from_buffer_copy(book[i].tobytes()), 4, 8) for i in range(len(book))]
positions = workloads.playout_positions(1_000_000)
offsets, moves = oracle.generate_moves_batch(positions)
sha = lambda a: hashlib.sha256(a.tobytes()).hexdigest()
out = {
    "book": {"name": "chess324-synth", "positions": len(book), "sha256": sha(book), "depth": 4, "perft": counts,
             "nodes": sum(counts)},
    "playouts": {"positions": len(positions), "sha256": sha(positions), "moves": int(len(moves)),
                 "offsets_sha256": sha(offsets), "moves_sha256": sha(moves)},
}
(HERE / "workload_fixtures.json").write_text(json.dumps(out) + "\n")
print({k: {a: b for a, b in v.items() if a != "perft"} for k, v in out.items()})

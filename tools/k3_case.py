"""1M-position move-list batch (BASELINE.json configs[4]) a few times, for ncu launch lists."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import workloads  # noqa: E402

import pabi_b200 as pb  # noqa: E402

e = pb.Engine(0)
positions = workloads.playout_positions(1_000_000)
for _ in range(3):
    offsets, moves = e.generate_moves_batch(positions)
    print(len(moves), round(e.last_timing.device_ms, 4), "ms device", e.last_timing.kernel_launches, "launches")

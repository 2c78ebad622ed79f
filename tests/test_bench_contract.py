"""bench.py's output contract: one JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert COMMON <= set(line) and line["impl"] == "reference"
    assert line["unit"] == "nodes/s" and line["value"] > 1e6 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    line = run_bench("--depth", "6", "--steps", "2", "--warmup", "3")
    assert COMMON <= set(line) and "impl" not in line
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["dtype"] == "u64"
    assert line["value"] > 1e9 and line["e2e"]["value"] > 1e9 and line["e2e"]["h2d_bytes_per_step"] == 112
    assert line["gpu_launches"] > 0
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 1e6 and c["sample"]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

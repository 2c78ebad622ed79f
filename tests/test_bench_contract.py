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
    assert line["cpu_baseline"]["single_thread"]["cores"] == 1
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
    assert r["bound"] == "int-issue" and r["unit"] == "Gwarp-inst/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.05 < r["frac"] < 1.0 and 0.05 < r["alu_pipe"]["frac"] < 1.0  # physical fractions
    h = r["hbm_model"]
    assert h["unit"] == "GB/s" and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 1e6 and c["sample"]
    assert c["single_thread"]["cores"] == 1 and c["single_thread"]["value"] > 1e6
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.gpu
def test_gpu_arm_reports_the_other_baseline_configs():
    """The headline line carries BASELINE.json configs 1, 2, 3 and 5 too, each checked and with its own figures."""
    line = run_bench("--steps", "2", "--warmup", "3", "--no-cpu")
    o = line["other_configs"]
    assert o["perft6_startpos"]["nodes"] == 119060324 and o["perft6_startpos"]["device_ms"] < 5
    assert o["perft_suite"]["vectors"] >= 78 and o["perft_suite"]["one_batch_per_root_depths"]["device_ms"] > 0
    assert o["perft4_book"]["positions"] == 1620 and o["perft4_book"]["nodes"] == 672560035
    m = o["move_lists_1m"]
    assert m["positions"] == 1000000 and m["moves"] == 27639035 and 0 < m["roofline"]["frac"] < 1.2
    assert line["roofline"]["instructions"]["exact_for_this_workload"] is True

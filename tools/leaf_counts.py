"""Writes profiles/leaf_kernel_counts.json -- what bench.py's issue roofline divides by -- from one GPU visit of
tools/gpu_round_r02.sh:

  gpurun_out/leaf_inst_<tag>.csv   ncu --metrics smsp__inst_executed.sum,... over the K2 launches of startpos perft(8)
                                   (one launch per perft since the ply above it is virtual; four while it was materialised)
  gpurun_out/leaf_<tag>.ncu-rep    ncu --set full capture of the depth-7 K2 launch (DRAM bytes, pipes, lanes)

    python tools/leaf_counts.py <tag>
"""
import csv
import hashlib
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
PARENTS, CHILDREN = 119_060_324, 3_195_901_860   # plies 6 and 7 of the start position

rows = list(csv.reader(open(ROOT / f"gpurun_out/leaf_inst_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
header = rows[hi]
per_launch = defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) >= len(header):
        rec = dict(zip(header, r))
        per_launch[int(rec["ID"])][rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
ids = sorted(per_launch)
assert len(ids) >= 2 and len(ids) % 2 == 0, "expected the leaf launches of two perft(8) calls (tools/quick_perft.py 8)"
LAUNCHES_PER_PERFT8 = len(ids) // 2   # 4 while ply 6 was materialised in chunks, 1 with the virtual ply
step = [per_launch[i] for i in ids[LAUNCHES_PER_PERFT8:]]  # the second (warm) perft(8)
tot = lambda k: sum(l[k] for l in step)

raw = subprocess.run(["ncu", "-i", str(ROOT / f"gpurun_out/leaf_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
cap = list(csv.reader(raw.splitlines()))
h, u, v = cap[0], cap[1], cap[2]
g = lambda k: float(v[h.index(k)].replace(",", ""))
scaled = lambda k: g(k) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u[h.index(k)], 1)
digest = hashlib.sha256()
for name in ("chess.cuh", "incremental.cuh", "kernels.cuh"):
    digest.update((ROOT / "pabi_b200" / "csrc" / name).read_bytes())
capture_parents = 4_865_609
out = {
    "source": f"ncu over the {LAUNCHES_PER_PERFT8} K2 launches of one startpos perft(8) on one B200 (gpurun_out/leaf_inst_{tag}.csv, "
              f"summarised in profiles/{tag}_leaf_inst.csv); per perft(8)",
    "kernel_source_digest": digest.hexdigest()[:16],
    "leaf_parents": PARENTS, "leaf_children": CHILDREN, "launches": LAUNCHES_PER_PERFT8,
    "warp_instructions": tot("smsp__inst_executed.sum"), "thread_instructions": tot("smsp__thread_inst_executed.sum"),
    "alu_instructions": tot("sm__inst_executed_pipe_alu.sum"), "fma_instructions": tot("sm__inst_executed_pipe_fma.sum"),
    "xu_instructions": tot("sm__inst_executed_pipe_xu.sum"), "lsu_instructions": tot("sm__inst_executed_pipe_lsu.sum"),
    "kernel_ms_under_ncu": tot("gpu__time_duration.sum") / 1e6, "sm_cycles_elapsed": tot("sm__cycles_elapsed.max"),
    "dram_bytes_per_parent": (scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")) / capture_parents,
    "capture": {
        "source": f"profiles/{tag}_leaf_kernel.txt: ncu --set full of the depth-7 launch ({capture_parents} parents, 119060324 children)",
        "kernel_ms": g("gpu__time_duration.sum") / (1e3 if u[h.index("gpu__time_duration.sum")] == "us" else 1e6 if u[h.index("gpu__time_duration.sum")] == "ns" else 1),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "thread_inst_per_warp_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "l1_hit_rate_pct": g("l1tex__t_sector_hit_rate.pct"),
        "shared_bank_conflict_pct": 100 * g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": g("launch__registers_per_thread"),
    },
}
(ROOT / "profiles" / "leaf_kernel_counts.json").write_text(json.dumps(out, indent=1) + "\n")
cycles = out["sm_cycles_elapsed"]
print(json.dumps({"warp_inst_per_child": out["warp_instructions"] / CHILDREN, "lanes": out["thread_instructions"] / out["warp_instructions"],
                  "ipc_per_scheduler": out["warp_instructions"] / (148 * 4 * cycles), "alu_pipe": out["alu_instructions"] / (148 * 4 * cycles * 0.5),
                  "ms_under_ncu": out["kernel_ms_under_ncu"], "capture": out["capture"]}, indent=1))

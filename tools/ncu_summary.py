"""Summarise one ncu report (first captured kernel) into a text file for profiles/.

    python tools/ncu_summary.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> <title> > profiles/<name>.txt
"""
import csv
import subprocess
import sys

rep, so, kern, title = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
print(f"# {title}")
print(f"# kernel: {v[h.index('Kernel Name')]}")
print(f"# report: {rep} (ncu --set full --clock-control none --import-source on)")
for k in KEYS:
    if k in h:
        print(f"{k:72s} {v[h.index(k)]:>20s} {u[h.index(k)]}")
print()
out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, so, kern, "40"], capture_output=True, text=True).stdout
print(out)

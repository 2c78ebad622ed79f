#!/bin/bash
# Copies the artefacts of one tools/gpu_round.sh visit from gpurun_out/ into profiles/ and refreshes
# profiles/leaf_kernel_traffic.json.  Usage: bash tools/save_round.sh <tag> "<mangled kernel substring>" "<title>"
set -eu
TAG=$1; KSUB=$2; TITLE=$3
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/bench_$TAG.json profiles/${TAG}_bench.json
[ -f gpurun_out/configs_$TAG.jsonl ] && cp gpurun_out/configs_$TAG.jsonl profiles/${TAG}_other_configs.jsonl
[ -f gpurun_out/suite_$TAG.jsonl ] && cp gpurun_out/suite_$TAG.jsonl profiles/${TAG}_suite_cases.jsonl
python tools/ncu_summary.py gpurun_out/leaf_$TAG.ncu-rep pabi_b200/libpabi_b200.so "$KSUB" "$TITLE" > profiles/${TAG}_leaf_kernel.txt
python - "$TAG" <<'PY'
import csv, json, subprocess, sys
tag = sys.argv[1]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/leaf_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
def val(k):
    return float(v[h.index(k)]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u[h.index(k)], 1)
g = lambda k: float(v[h.index(k)])
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
parents = 4865609  # the depth-7 launch of bench.py --depth 7
d = {"source": f"profiles/{tag}_leaf_kernel.txt: ncu --set full capture of k_tile<TILE_LEAF> at depth 7 ({parents} parents per launch)",
     "dram_bytes_read": rd, "dram_bytes_write": wr, "parents_in_capture": parents, "dram_bytes_per_parent": (rd + wr) / parents,
     "alu_pipe_pct_of_peak": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
     "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     "warp_instructions": g("smsp__inst_executed.sum"), "children_in_capture": 119060324,
     "thread_inst_per_warp_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "l1_hit_rate_pct": g("l1tex__t_sector_hit_rate.pct"),
     "shared_bank_conflict_pct": 100 * g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "kernel_ms_in_capture": g("gpu__time_duration.sum"),
     "note": "k_tile<TILE_LEAF> is bound by the integer ALU pipe (64-bit bitboard logic on a 32-bit datapath), not by HBM"}
json.dump(d, open("profiles/leaf_kernel_traffic.json", "w"), indent=1)
print({k: round(x, 2) if isinstance(x, float) else x for k, x in d.items() if k != "source" and k != "note"})
PY

#!/bin/bash
# One GPU visit for the kernels either side of the leaf kernel: full ncu captures of K1 (k_tile<TILE_EXPAND>, a
# ply-5-to-6 chunk of startpos perft(8)) and K3 (k_moves_warp on the 1 M playout positions), each after the same
# command has exited 0 without ncu.  Usage (under gpurun): bash tools/gpu_side_kernels.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
K1="python bench.py --depth 8 --steps 1 --warmup 1 --no-cpu"
$K1 > $OUT/plain_k1_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tile<\(pabi::TileMode\)1' -s 2 -c 1 -f -o $OUT/k1_$TAG $K1 > $OUT/ncu_k1_$TAG.log 2>&1
echo "K1 capture rc=$?"; tail -3 $OUT/ncu_k1_$TAG.log
K3="python tools/k3_case.py"
$K3 > $OUT/plain_k3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_moves_warp -s 1 -c 1 -f -o $OUT/k3_$TAG $K3 > $OUT/ncu_k3_$TAG.log 2>&1
echo "K3 capture rc=$?"; tail -3 $OUT/ncu_k3_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_k3_$TAG.csv $K3 > $OUT/ncu_k3l_$TAG.log 2>&1
echo "K3 launch list rc=$?"
ls -la $OUT | tail -12

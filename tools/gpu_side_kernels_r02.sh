#!/bin/bash
# Full ncu captures of the kernels either side of K2: K1 (k_tile<TILE_EXPAND>, a ply-5-to-6 chunk of startpos perft(8)) and
# the fused K3 (k_moves_fused on the 1 M playout positions, device-resident records), each after the same command has
# exited 0 without ncu.  Usage (under gpurun): bash tools/gpu_side_kernels_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
K1="python tools/quick_perft.py 8"
$K1 > $OUT/plain_k1_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tile<\(pabi::TileMode\)1' -s 3 -c 1 -f -o $OUT/k1_$TAG $K1 > $OUT/ncu_k1_$TAG.log 2>&1
echo "K1 capture rc=$?"; tail -2 $OUT/ncu_k1_$TAG.log
K3="python tools/list_device_check.py"
$K3 > $OUT/plain_k3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_moves_fused -s 2 -c 1 -f -o $OUT/k3_$TAG $K3 > $OUT/ncu_k3_$TAG.log 2>&1
echo "K3 capture rc=$?"; tail -2 $OUT/ncu_k3_$TAG.log
cat $OUT/plain_k3_$TAG.log | tail -1
ls -la $OUT | tail -6

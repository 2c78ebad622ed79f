#!/bin/bash
# Full ncu captures of the kernels either side of K2: K1 (k_tile<TILE_EXPAND_VIRTUAL>, the virtual ply 5-to-6 of startpos
# perft(8); with PABI_NO_VIRTUAL_PLY=1 in the environment: k_tile<TILE_EXPAND>, a materialised ply-5-to-6 chunk) and the fused K3 (k_moves_fused on the 1 M playout positions, device-resident records), each after the same command has
# exited 0 without ncu.  Usage (under gpurun): bash tools/gpu_side_kernels_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
K1="python tools/quick_perft.py 8"
$K1 > $OUT/plain_k1_$TAG.log 2>&1 &&
MODE=3; SKIP=1
if [ "${PABI_NO_VIRTUAL_PLY:-0}" = "1" ]; then MODE=1; SKIP=3; fi
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_tile<\(pabi::TileMode\)$MODE" -s $SKIP -c 1 -f -o $OUT/k1_$TAG $K1 > $OUT/ncu_k1_$TAG.log 2>&1
echo "K1 capture rc=$?"; tail -2 $OUT/ncu_k1_$TAG.log
K3="python tools/list_device_check.py"
$K3 > $OUT/plain_k3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_moves_fused -s 2 -c 1 -f -o $OUT/k3_$TAG $K3 > $OUT/ncu_k3_$TAG.log 2>&1
echo "K3 capture rc=$?"; tail -2 $OUT/ncu_k3_$TAG.log
cat $OUT/plain_k3_$TAG.log | tail -1
ls -la $OUT | tail -6

#!/bin/bash
# Copies the artefacts of one tools/gpu_round_r02.sh (+ gpu_side_kernels_r02.sh) visit from gpurun_out/ into profiles/ and
# refreshes profiles/leaf_kernel_counts.json (what bench.py's issue roofline divides by).  Usage: bash tools/save_round_r02.sh <tag>
set -eu
TAG=$1
K2="k_tileILNS_8TileModeE0ELNS_10AccountingE0ENS_10TileConfigILi1024ELi512ELi14080ELi1ELi2"
K1="k_tileILNS_8TileModeE3ELNS_10AccountingE0ENS_10TileConfigILi1024ELi512ELi14080ELi1ELi2"
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/leaf_inst_$TAG.csv profiles/${TAG}_leaf_inst.csv
cp gpurun_out/bench_$TAG.json profiles/${TAG}_bench.json
python tools/leaf_counts.py $TAG > /dev/null
python tools/ncu_summary.py gpurun_out/leaf_$TAG.ncu-rep pabi_b200/libpabi_b200.so "$K2" "K2 k_tile<TILE_LEAF, SINGLE_ROOT> of round 2 ($TAG), depth-7 launch: 4 865 609 parents, 119 060 324 children; bench.py --depth 7" > profiles/${TAG}_leaf_kernel.txt
[ -f gpurun_out/k1_$TAG.ncu-rep ] && python tools/ncu_summary.py gpurun_out/k1_$TAG.ncu-rep pabi_b200/libpabi_b200.so "$K1" "K1 k_tile<TILE_EXPAND_VIRTUAL, SINGLE_ROOT> of round 2 ($TAG): the virtual ply 5-to-6 of startpos perft(8), 4 865 609 parents -> 119 060 324 eight-byte entries" > profiles/${TAG}_k1_expand.txt
[ -f gpurun_out/k3_$TAG.ncu-rep ] && python tools/ncu_summary.py gpurun_out/k3_$TAG.ncu-rep pabi_b200/libpabi_b200.so "k_moves_fused" "K3 k_moves_fused of round 2 ($TAG): ordered move lists of the 1 M playout positions, device-resident records, one launch" > profiles/${TAG}_k3_moves_fused.txt
wc -l profiles/${TAG}_*.txt

#!/bin/bash
# One GPU visit: parity tests, the bench line, an ncu launch list and one full capture of the
# dominant kernel.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
LIST="python bench.py --steps 2 --warmup 1 --no-cpu"
$LIST > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $LIST > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --depth 7 --steps 1 --warmup 1 --no-cpu"
$SMALL > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 5 -c 1 -f -o $OUT/leaf_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la $OUT

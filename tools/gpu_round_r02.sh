#!/bin/bash
# One GPU visit of round 2: parity tests, the bench line, the ncu launch list of the bench command, one
# `--set full` capture of the leaf kernel (depth-7 launch) and the per-launch instruction counts of the
# depth-8 leaf launches (what bench.py's issue roofline divides by).  Each ncu run directly follows the
# same command exiting 0 without ncu.  Usage (under gpurun): bash tools/gpu_round_r02.sh <tag> [skip-tests]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log
fi
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
LIST="python bench.py --steps 2 --warmup 1 --no-cpu --no-extra"
$LIST > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $LIST > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
SMALL="python bench.py --depth 7 --steps 1 --warmup 1 --no-cpu --no-extra"
$SMALL > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tile<\(pabi::TileMode\)0' -s 1 -c 1 -f -o $OUT/leaf_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
# instruction counts of every leaf launch of one perft(8) (second call of the process: warm)
D8="python tools/quick_perft.py 8"
$D8 > $OUT/plain3_$TAG.log 2>&1 &&
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum \
  --clock-control none --kernel-name-base demangled -k 'regex:k_tile<\(pabi::TileMode\)0' -c 2 --csv --log-file $OUT/leaf_inst_$TAG.csv $D8 > $OUT/ncu_inst_$TAG.log 2>&1
echo "inst count rc=$?"
ls -la $OUT | tail -12

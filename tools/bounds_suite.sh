#!/bin/bash
# compute-sanitizer is closed on this GPU pool, so memory safety of the kernels' index arithmetic (atomically reserved pool
# slots, scanned child offsets, chunked frontiers) is checked with the -DPABI_BOUNDS build instead: the whole GPU suite runs
# against build/variants/bounds.so (every record / pool index tested, see chess.cuh) and the violation flag is read at the end.
# Usage (under gpurun, after `make -C pabi_b200/csrc bounds` here): bash tools/bounds_suite.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
cp pabi_b200/libpabi_b200.so /tmp/base.so
cp build/variants/bounds.so pabi_b200/libpabi_b200.so
python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_bounds_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_bounds_$TAG.log
python - > $OUT/bounds_$TAG.json <<'PY'
import json
import pabi_b200 as pb
e = pb.Engine(0)
small = pb.Engine(0, 20000)
out = {"build": "PABI_BOUNDS", "before": e.bounds_violation()}
start = pb.Position.starting()
kiwi = pb.Position.try_from("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
dense = pb.Position.try_from("knq5/ppQ1qQQ1/2q4q/5Q2/1Q1Q3q/2qQ3q/Q4qPP/q1Q3NK w - - 0 1")
out["start_d7"] = e.perft(start, 7) == 3195901860
out["dense_d4"] = e.perft(dense, 4) == 169711264           # many moves per parent: pool pressure, halved ranges
out["small_kiwi_d5"] = small.perft(kiwi, 5) == 193690690    # chunked frontiers
out["hashed_d7"] = e.perft_hashed(start, 7) == 3195901860
out["batch"] = e.perft_batch([start, kiwi, dense], 4).tolist() == [197281, 4085603, 169711264]
out["violation_line_engine"] = e.bounds_violation()
out["violation_line_small"] = small.bounds_violation()
print(json.dumps(out))
PY
cat $OUT/bounds_$TAG.json; tail -3 $OUT/pytest_bounds_$TAG.log
cp /tmp/base.so pabi_b200/libpabi_b200.so

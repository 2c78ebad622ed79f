#!/bin/bash
# Development helper (under gpurun): exactness gate + timings of the small configs with and without speculative plies.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-fp}
python tools/variant_gate.py 100 > $OUT/gate_$TAG.json 2>&1; echo "gate rc=$?"; cat $OUT/gate_$TAG.json
for mode in 0 1; do
  echo "== PABI_EXACT_LEVELS=$mode"
  PABI_EXACT_LEVELS=$mode python tools/quick_perft.py 4,5,6,7 2>&1 | grep -v '"depth": 7' 
  PABI_EXACT_LEVELS=$mode python tools/suite_cases.py 2>&1 | tail -2
done | tee $OUT/fast_$TAG.log
python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/pytest_$TAG.log

#!/bin/bash
# Development helper (under gpurun): gate a library build from build/variants/ (exact counts, perft(8) below a
# threshold), then the whole GPU suite on it, then the other tile geometries.  Usage: bash tools/variant_gate.sh <name> <ms>
set -u
V=${1:-overlap}; LIMIT=${2:-41.0}
OUT=gpurun_out; mkdir -p $OUT
cp pabi_b200/libpabi_b200.so /tmp/base.so
cp build/variants/$V.so pabi_b200/libpabi_b200.so
python tools/variant_gate.py $LIMIT > $OUT/gate_$V.json 2>&1; rc=$?
cat $OUT/gate_$V.json; echo "gate rc=$rc"
if [ $rc -eq 0 ]; then
  python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_$V.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$V.log
  tail -4 $OUT/pytest_$V.log
fi
for t in 4 5 1; do echo "== tile variant $t"; PABI_TILE_VARIANT=$t python tools/quick_perft.py 8 2>&1 | tail -1; done | tee $OUT/sweep_$V.log
cp /tmp/base.so pabi_b200/libpabi_b200.so

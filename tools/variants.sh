#!/bin/bash
# Development helper (under gpurun): time startpos perft(7)/(8) with every library build under build/variants/
# against the in-tree library.  Usage: bash tools/variants.sh [depths]
set -u
D=${1:-7,8}
OUT=gpurun_out
mkdir -p $OUT
cp pabi_b200/libpabi_b200.so /tmp/base.so
echo "== base" | tee $OUT/variants.log
python tools/quick_perft.py $D 2>&1 | tee -a $OUT/variants.log
for v in build/variants/*.so; do
  echo "== $v" | tee -a $OUT/variants.log
  cp $v pabi_b200/libpabi_b200.so
  timeout 120 python tools/quick_perft.py $D 2>&1 | tee -a $OUT/variants.log
done
cp /tmp/base.so pabi_b200/libpabi_b200.so
echo "== base again" | tee -a $OUT/variants.log
python tools/quick_perft.py $D 2>&1 | tee -a $OUT/variants.log

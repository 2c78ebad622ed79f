#!/bin/bash
# Development helper (under gpurun): run a command with the in-tree library and with every build/variants/<name>.so
# named on the command line swapped in.  Usage: bash tools/variants_cmd.sh "<command>" name1 name2 ...
set -u
CMD=$1; shift
cp pabi_b200/libpabi_b200.so /tmp/base.so
echo "== base"; eval "$CMD"
for v in "$@"; do
  echo "== $v"; cp build/variants/$v.so pabi_b200/libpabi_b200.so; eval "timeout 300 $CMD"
done
cp /tmp/base.so pabi_b200/libpabi_b200.so

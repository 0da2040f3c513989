#!/bin/bash
# GPU session: where partitioning starts to pay (genome size sweep), L2 hints with and without partitioning.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2e}
for n in 1e8 2e8 4e8 8e8 1.6e9; do
  timeout 600 python tools/part_sweep.py $n 5e7 packed4,plain 0,6,8 4,5 3,27 > $OUT/${TAG}_size_$n.log 2>&1; echo "size $n rc=$?"; grep Gq $OUT/${TAG}_size_$n.log | grep "mut 0"
done
timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 0,8 4,5 3,15,27 > $OUT/${TAG}_size_3.1e9.log 2>&1; echo "size 3.1e9 rc=$?"; grep Gq $OUT/${TAG}_size_3.1e9.log | grep "mut 0"
cat $OUT/part_sweep_*.json > /dev/null

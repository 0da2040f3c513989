#!/bin/bash
# GPU session: L2 eviction hints on the partitioned path at c3; ncu of the partition passes.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2d}
timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 8,10 4,5 3,7,19,11,27 > $OUT/${TAG}_part_c3_hints.log 2>&1; echo "part c3 hints rc=$?"; grep Gq $OUT/${TAG}_part_c3_hints.log | grep "mut 0"
cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_hints.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"part_scatter_staged|part_unpermute_flat|part_hist" -c 3 \
  -f -o $OUT/${TAG}_c3_part_passes python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 3 > $OUT/${TAG}_ncu_passes.log 2>&1; echo "ncu rc=$?"

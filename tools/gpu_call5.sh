#!/bin/bash
# GPU session: partition parity test, stage breakdown of the partitioned path (direct against staged scatter) at c2 / c3.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1x}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned" > $OUT/${TAG}_pytest_part.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_part.log
for sc in 0 1; do
  SAPLING_B200_PART_SCATTER=$sc timeout 600 python tools/part_sweep.py 1e8 5e7 plain,packed4 0,4,8,11 4,5 3 > $OUT/${TAG}_part_c2_s$sc.log 2>&1; echo "part c2 scatter=$sc rc=$?"; grep Gq $OUT/${TAG}_part_c2_s$sc.log
  cp $OUT/part_sweep_100000000.json $OUT/${TAG}_part_sweep_c2_s$sc.json
done
for sc in 0 1; do
  SAPLING_B200_PART_SCATTER=$sc timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,plain 0,6,8,11 4 3 > $OUT/${TAG}_part_c3_s$sc.log 2>&1; echo "part c3 scatter=$sc rc=$?"; grep Gq $OUT/${TAG}_part_c3_s$sc.log
  cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_s$sc.json
done

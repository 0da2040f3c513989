#!/bin/bash
# GPU session: in-order software-pipelined kernel (SAPLING_B200_ORDERED_PIPE) on the partitioned path, overlapping rank lines.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2a}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned" > $OUT/${TAG}_pytest_part.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_part.log
for op in 1 0; do
  SAPLING_B200_ORDERED_PIPE=$op timeout 600 python tools/part_sweep.py 1e8 5e7 plain,packed4 0,8,10 4,5 3 > $OUT/${TAG}_part_c2_o$op.log 2>&1; echo "part c2 ordered_pipe=$op rc=$?"; grep Gq $OUT/${TAG}_part_c2_o$op.log | grep "mut 0"
  cp $OUT/part_sweep_100000000.json $OUT/${TAG}_part_sweep_c2_o$op.json
done
SAPLING_B200_ORDERED_PIPE=1 timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,packed3,inline 0,8,10,11 3,4,5 3 > $OUT/${TAG}_part_c3_o1.log 2>&1; echo "part c3 ordered_pipe=1 rc=$?"; grep Gq $OUT/${TAG}_part_c3_o1.log
cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_o1.json

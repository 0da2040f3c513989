#!/bin/bash
# GPU session: in-order tile schedule (SAPLING_B200_PART_TILES) and anchor-line L1 prefetch (hint bit 16) on the partitioned path.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1z}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned" > $OUT/${TAG}_pytest_part.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_part.log
for tl in 1 0; do
  SAPLING_B200_PART_TILES=$tl timeout 600 python tools/part_sweep.py 1e8 5e7 plain,packed4 0,6,8,10 4,5 3,19 > $OUT/${TAG}_part_c2_t$tl.log 2>&1; echo "part c2 tiles=$tl rc=$?"; grep Gq $OUT/${TAG}_part_c2_t$tl.log | grep "mut 0"
  cp $OUT/part_sweep_100000000.json $OUT/${TAG}_part_sweep_c2_t$tl.json
done
SAPLING_B200_PART_TILES=1 timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,plain 0,6,8,10 4,5 3,19 > $OUT/${TAG}_part_c3_t1.log 2>&1; echo "part c3 tiles=1 rc=$?"; grep Gq $OUT/${TAG}_part_c3_t1.log
cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_t1.json

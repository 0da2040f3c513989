#!/bin/bash
# GPU session: full parity suite with the lean replay and the flat un-permute; A/B of both at c2 / c3.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
for ln in 1 0; do
  SAPLING_B200_LEAN=$ln timeout 600 python tools/part_sweep.py 1e8 5e7 plain,packed4 0,6,8 4,5 3 > $OUT/${TAG}_part_c2_l$ln.log 2>&1; echo "part c2 lean=$ln rc=$?"; grep Gq $OUT/${TAG}_part_c2_l$ln.log
  cp $OUT/part_sweep_100000000.json $OUT/${TAG}_part_sweep_c2_l$ln.json
done
for ln in 1 0; do
  SAPLING_B200_LEAN=$ln timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,inline 0,8,9,10 4,5 3 > $OUT/${TAG}_part_c3_l$ln.log 2>&1; echo "part c3 lean=$ln rc=$?"; grep Gq $OUT/${TAG}_part_c3_l$ln.log | grep "mut 0"
  cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_l$ln.json
done
SAPLING_B200_PART_UNPERMUTE=0 timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 8,10 5 3 > $OUT/${TAG}_part_c3_u0.log 2>&1; echo "part c3 unpermute=0 rc=$?"; grep Gq $OUT/${TAG}_part_c3_u0.log | grep "mut 0"

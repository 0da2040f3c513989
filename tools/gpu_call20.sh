#!/bin/bash
# GPU session: parked binarySearch tails (kMode 5) against the plain in-order kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2o}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partition or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_sel.log
for pk in 1 0; do
  SAPLING_B200_PARK=$pk timeout 600 python tools/part_sweep.py 1e8 5e7 packed4 5 4,5,6 27 > $OUT/${TAG}_c2_pk$pk.log 2>&1; echo "c2 park=$pk rc=$?"; grep Gq $OUT/${TAG}_c2_pk$pk.log
done
SAPLING_B200_PARK=1 timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 4,5 27 > $OUT/${TAG}_c3_pk1.log 2>&1; echo "c3 park=1 rc=$?"; grep Gq $OUT/${TAG}_c3_pk1.log

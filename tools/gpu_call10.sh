#!/bin/bash
# GPU session: variants parity test; anchor line in shared memory (SAPLING_B200_LINE_SMEM) A/B at c2 / c3; 8192-query chunks.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2c}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partitioned or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_sel.log
for ls in 1 0; do
  SAPLING_B200_LINE_SMEM=$ls timeout 600 python tools/part_sweep.py 1e8 5e7 packed4,packed3 0,6 4,5 3 > $OUT/${TAG}_part_c2_ls$ls.log 2>&1; echo "part c2 line_smem=$ls rc=$?"; grep Gq $OUT/${TAG}_part_c2_ls$ls.log | grep "mut 0"
  cp $OUT/part_sweep_100000000.json $OUT/${TAG}_part_sweep_c2_ls$ls.json
done
for ls in 1 0; do
  SAPLING_B200_LINE_SMEM=$ls timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,packed3 0,8,10 3,4,5 3 > $OUT/${TAG}_part_c3_ls$ls.log 2>&1; echo "part c3 line_smem=$ls rc=$?"; grep Gq $OUT/${TAG}_part_c3_ls$ls.log | grep "mut 0"
  cp $OUT/part_sweep_3100000000.json $OUT/${TAG}_part_sweep_c3_ls$ls.json
done

#!/bin/bash
# GPU session: partition parity test, partition sweep at c2 / c3, launch list of one partitioned run.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1w}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned" > $OUT/${TAG}_pytest_part.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/${TAG}_pytest_part.log
timeout 600 python tools/part_sweep.py 1e8 5e7 plain,packed4 0,3,4,5,6,8 4,5 3 > $OUT/${TAG}_part_c2.log 2>&1; echo "part c2 rc=$?"; grep Gq $OUT/${TAG}_part_c2.log
timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4,inline 0,6,8,10,11 4 3 > $OUT/${TAG}_part_c3.log 2>&1; echo "part c3 rc=$?"; grep Gq $OUT/${TAG}_part_c3.log
SAPLING_B200_PART_MIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
tail -3 $OUT/${TAG}_ncu_bench.log

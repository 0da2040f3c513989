#!/bin/bash
# GPU session: quick parity subset + default-path timings at c2 / c3 (A/B against the previous call's numbers).
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2n}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partition or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_sel.log
timeout 600 python tools/part_sweep.py 1e8 5e7 packed4 0,5 5 27 > $OUT/${TAG}_c2.log 2>&1; echo "c2 rc=$?"; grep Gq $OUT/${TAG}_c2.log
timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 27 > $OUT/${TAG}_c3.log 2>&1; echo "c3 rc=$?"; grep Gq $OUT/${TAG}_c3.log

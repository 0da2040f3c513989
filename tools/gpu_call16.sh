#!/bin/bash
# GPU session: division slow path avoided (x == xlo), tiles claimed four at a time, 6 blocks/SM variant.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2i}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partition or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_sel.log
timeout 600 python tools/part_sweep.py 1e8 5e7 packed4 0,5 4,5,6 27 > $OUT/${TAG}_c2.log 2>&1; echo "c2 rc=$?"; grep Gq $OUT/${TAG}_c2.log
timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5,6 27 > $OUT/${TAG}_c3.log 2>&1; echo "c3 rc=$?"; grep Gq $OUT/${TAG}_c3.log

#!/bin/bash
# GPU session: flat replay (kmer_replay_flat) against the lean one, table-based staged scatter, lane-group un-permute.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2g}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partition or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_sel.log
for fl in 1 0; do
  bps=4,5; [ $fl = 1 ] && bps=4,5,6
  SAPLING_B200_FLAT=$fl timeout 600 python tools/part_sweep.py 1e8 5e7 packed4 0,6 $bps 27 > $OUT/${TAG}_c2_flat$fl.log 2>&1; echo "c2 flat=$fl rc=$?"; grep Gq $OUT/${TAG}_c2_flat$fl.log
done
for fl in 1 0; do
  bps=4,5; [ $fl = 1 ] && bps=4,5,6
  SAPLING_B200_FLAT=$fl timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 0,10 $bps 27 > $OUT/${TAG}_c3_flat$fl.log 2>&1; echo "c3 flat=$fl rc=$?"; grep Gq $OUT/${TAG}_c3_flat$fl.log
done
SAPLING_B200_PART_UNPERMUTE=1 timeout 600 python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 27 > $OUT/${TAG}_c3_unp1.log 2>&1; echo "c3 unpermute=flat rc=$?"; grep Gq $OUT/${TAG}_c3_unp1.log

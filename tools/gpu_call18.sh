#!/bin/bash
# GPU session: slot inside the k-mer word, coalesced row scan, un-permute with 8 runs in flight per lane group.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2m}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or partition or rank_line" > $OUT/${TAG}_pytest_sel.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_sel.log
for sk in 1 0; do
  SAPLING_B200_SLOT_IN_KMER=$sk timeout 600 python tools/part_sweep.py 1e8 5e7 packed4 5 5 27 > $OUT/${TAG}_c2_sk$sk.log 2>&1; echo "c2 slot_in_kmer=$sk rc=$?"; grep Gq $OUT/${TAG}_c2_sk$sk.log
  SAPLING_B200_SLOT_IN_KMER=$sk timeout 900 python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 27 > $OUT/${TAG}_c3_sk$sk.log 2>&1; echo "c3 slot_in_kmer=$sk rc=$?"; grep Gq $OUT/${TAG}_c3_sk$sk.log
done

#!/bin/bash
# GPU session: ncu --set full of the query kernel inside a partitioned c3 step (bits=8, tiling rank lines).
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1y}
export SAPLING_B200_PART_BITS=8 SAPLING_B200_PACKED_SHIFT=4
timeout 800 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 \
  -f -o $OUT/${TAG}_c3_part_query python bench.py --workload c3 --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 \
  > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu-full rc=$?"
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-600
ls -la $OUT/*.ncu-rep

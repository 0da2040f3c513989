#!/bin/bash
# GPU session: full state of the tree -- parity tests, bench c2 / c3, reference arm, ncu launch list, ncu --set full of the
# in-order query kernel at c2 and c3 and of the partition passes.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2k}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_box.txt 2>&1; nproc >> $OUT/${TAG}_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench c2 rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json
timeout 900 python bench.py --workload c3 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-300 $OUT/${TAG}_bench_c3.json
timeout 600 python bench.py --impl reference --steps 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 \
  > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 \
  -f -o $OUT/${TAG}_c2_query python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 \
  > $OUT/${TAG}_ncu_full_c2.log 2>&1; echo "ncu-full c2 rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 \
  -f -o $OUT/${TAG}_c3_query python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 27 \
  > $OUT/${TAG}_ncu_full_c3.log 2>&1; echo "ncu-full c3 rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"part_scatter_staged|part_unpermute|part_hist" -s 9 -c 3 \
  -f -o $OUT/${TAG}_c3_passes python tools/part_sweep.py 3.1e9 2.5e8 packed4 10 5 27 \
  > $OUT/${TAG}_ncu_passes_c3.log 2>&1; echo "ncu-passes c3 rc=$?"
ls -la $OUT | tail -14

#!/bin/bash
# GPU session: parity tests, layout A/B at c2 and c3, bench, ncu.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1u}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_box.txt 2>&1; nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python tools/layouts.py 1e8 5e7 plain,packed3,packed4 3,7 > $OUT/${TAG}_layouts_c2.log 2>&1; echo "layouts c2 rc=$?"
timeout 600 python tools/layouts.py 4e8 5e7 plain,packed3,packed4 3 > $OUT/${TAG}_layouts_400m.log 2>&1; echo "layouts 400m rc=$?"
timeout 900 python tools/layouts.py 3.1e9 2.5e8 inline,packed3,packed4 3 > $OUT/${TAG}_layouts_c3.log 2>&1; echo "layouts c3 rc=$?"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 900 python bench.py --workload c3 --steps 5 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; cat $OUT/${TAG}_bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 -f -o $OUT/${TAG}_query \
  python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu-full rc=$?"
SAPLING_B200_PACKED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 -f -o $OUT/${TAG}_query_plain \
  python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > $OUT/${TAG}_ncu_full_plain.log 2>&1; echo "ncu-full-plain rc=$?"
ls -la $OUT

#!/bin/bash
# One GPU session: parity tests, experiments, bench, ncu launch list + full capture of the query kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [steps...]   steps: tests variants gather bench ncu
set -u
TAG=${1:-run}; shift || true
STEPS=${*:-tests variants bench ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_box.txt 2>&1
nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt
for s in $STEPS; do
  case $s in
    tests)   timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?";;
    variants) timeout 600 python tools/gpu_experiments.py variants > $OUT/${TAG}_variants.log 2>&1; echo "variants rc=$?";;
    footprint) timeout 600 python tools/gpu_experiments.py footprint > $OUT/${TAG}_footprint.log 2>&1; echo "footprint rc=$?";;
    tlb) timeout 600 python tools/gpu_experiments.py tlb > $OUT/${TAG}_tlb.log 2>&1; echo "tlb rc=$?";;
    hints|sizes|kernels|pin|nbsweep) timeout 900 python tools/gpu_experiments.py $s > $OUT/${TAG}_$s.log 2>&1; echo "$s rc=$?";;
    gather)  timeout 600 python tools/gpu_experiments.py gather > $OUT/${TAG}_gather.log 2>&1; echo "gather rc=$?";;
    bench)   timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?";;
    benchref) timeout 900 python bench.py --impl reference --steps 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "benchref rc=$?";;
    ncu)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 \
        > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query -s 3 -c 1 \
        -f -o $OUT/${TAG}_query python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 \
        > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu-full rc=$?";;
  esac
done
ls -la $OUT

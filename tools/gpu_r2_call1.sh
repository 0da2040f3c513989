#!/bin/bash
# round 2, call 1: box facts, the c3 bench (both arms) with the reference pinned through ref_from_parts, new tests
mkdir -p gpurun_out
{
  echo "== box"; nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  nvidia-smi topo -m 2>/dev/null | head -20
} > gpurun_out/s1_box.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "probe_count" > gpurun_out/s1_pytest_probe.log 2>&1
tail -3 gpurun_out/s1_pytest_probe.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/s1_bench_c3.json 2> gpurun_out/s1_bench_c3.log
tail -5 gpurun_out/s1_bench_c3.log; head -c 600 gpurun_out/s1_bench_c3.json
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/s1_bench_c3_ref.json 2> gpurun_out/s1_bench_c3_ref.log
tail -5 gpurun_out/s1_bench_c3_ref.log; head -c 600 gpurun_out/s1_bench_c3_ref.json

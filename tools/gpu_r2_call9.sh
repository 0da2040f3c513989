#!/bin/bash
# round 2, call 9: persistent scatter; what travels a tile ahead x occupancy once more (the kernel is lighter now)
mkdir -p gpurun_out
T=s9
timeout 900 python -m pytest tests -x -q -m gpu -k "partition or golden or chr3 or stray" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
for occ in 4 5; do for qv in 0 2 4; do
  SAPLING_B200_TUNE="occ=$occ,qv=$qv" timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_c3_occ${occ}_qv${qv}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c3_occ${occ}_qv${qv}.json')); print('c3 occ $occ qv $qv', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9))"
done; done
for qv in 0 2 4; do
  SAPLING_B200_TUNE="occ=4,qv=$qv" timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_c2_occ4_qv${qv}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c2_occ4_qv${qv}.json')); print('c2 occ 4 qv $qv', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_ -s 16 -c 8 -o gpurun_out/${T}_c3_passes -f python bench.py --steps 2 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_passes.log 2>&1; tail -1 gpurun_out/${T}_ncu_passes.log

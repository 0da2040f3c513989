#!/bin/bash
# round 2, call 11: the in-order kernel with its pipeline in shared memory (cp.async rings) against the register pipeline
mkdir -p gpurun_out
T=s11
SAPLING_B200_TUNE="occ=14" timeout 900 python -m pytest tests -x -q -m gpu -k "partition or golden or chr3 or stray or large or wide" > gpurun_out/${T}_pytest_ring.log 2>&1; echo "pytest ring rc=$?"; tail -3 gpurun_out/${T}_pytest_ring.log
for w in c3 c2; do for occ in 4 14 15 13; do
  SAPLING_B200_TUNE="occ=$occ" timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_${w}_occ${occ}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_occ${occ}.json')); print('$w occ $occ', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'selfcheck', d.get('self_check'))"
done; done
for occ in 14 15; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ring -s 3 -c 1 -o gpurun_out/${T}_c3_ring_occ$occ -f env SAPLING_B200_TUNE="occ=$occ" python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3_$occ.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3_$occ.log | head -c 200; echo
done

#!/bin/bash
# round 2, call 12: ring kernel (cp.async without cache hints) against the register pipeline; the bit-stream upload format
mkdir -p gpurun_out
T=s12
timeout 900 python -m pytest tests -x -q -m gpu -k "bit_stream or partitioned_large or replicas" > gpurun_out/${T}_pytest_bits.log 2>&1; echo "pytest bits rc=$?"; tail -3 gpurun_out/${T}_pytest_bits.log
for w in c3 c2; do for occ in 4 14 15 13; do
  SAPLING_B200_TUNE="occ=$occ" timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_${w}_occ${occ}.log | tail -1 > gpurun_out/${T}_${w}_occ${occ}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_${w}_occ${occ}.json')); print('$w occ $occ', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'selfcheck', d.get('self_check'))" || tail -5 gpurun_out/${T}_${w}_occ${occ}.log
done; done
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 5 2>/dev/null | tail -1 > gpurun_out/${T}_c3_e2e.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_c3_e2e.json')); e=d['e2e']; print('e2e bits %.3f bytes %.3f int64 %.3f equal %s' % (e['value']/1e9, e['byte_api']['value']/1e9, e['int64_api']['value']/1e9, e['answers_equal_device_path']))"
for occ in 14 15; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ring -s 3 -c 1 -o gpurun_out/${T}_c3_ring_occ$occ -f env SAPLING_B200_TUNE="occ=$occ" python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3_$occ.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3_$occ.log | head -c 200; echo
done

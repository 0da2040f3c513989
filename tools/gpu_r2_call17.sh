#!/bin/bash
# round 2, call 17: answers stored a tile after their rev[rank] was requested (occ=14) against stored in place
mkdir -p gpurun_out
T=s17
SAPLING_B200_TUNE="occ=14" timeout 900 python -m pytest tests -x -q -m gpu -k "partitioned_large or golden or chr3" > gpurun_out/${T}_pytest_defer.log 2>&1; echo "pytest late rc=$?"; tail -2 gpurun_out/${T}_pytest_defer.log
run() {  # workload, tune
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'sustained %.3f ms' % d['sustained']['ms_per_step'], 'bits', d['roofline']['partition_bits'], 'ok' if d['self_check']['matching']==d['self_check']['of'] else d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "occ=14" "" "occ=14" ; do run c3 "$tune"; done
for tune in "" "occ=14" ; do run c2 "$tune"; done
for tune in "" "occ=14"; do run c4 "$tune"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c3_defer -f env SAPLING_B200_TUNE="occ=14" python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3.log | head -c 200; echo

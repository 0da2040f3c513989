#!/bin/bash
# round 2, call 19: k = 31 (wide model, every match decided by the genome) against the slice count
mkdir -p gpurun_out
T=s19
run() {  # workload, tune, extra args
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 $3 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 $3 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'bits', d['roofline']['partition_bits'], d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "part_bits=10" "part_bits=11" "part_bits=8"; do run c4 "$tune" "--k 31"; done
for tune in "" "part_bits=10"; do run c3 "$tune" "--k 31"; done
for tune in "" "part_bits=10"; do run c4 "$tune" "--k 32"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c4_k31 -f python bench.py --workload c4 --k 31 --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1; tail -1 gpurun_out/${T}_ncu.log | head -c 200; echo

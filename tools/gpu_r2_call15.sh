#!/bin/bash
# round 2, call 15: un-permute with lane groups at the mean run length and the first answers of a batch requested together;
# at most 512 slices by default; k-mer ring as the one variant
mkdir -p gpurun_out
T=s15
timeout 900 python -m pytest tests -x -q -m gpu -k "partition or golden or chr3 or bit_stream or variants or stray" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
run() {  # workload, tune
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'sustained %.3f ms' % d['sustained']['ms_per_step'], 'bits', d['roofline']['partition_bits'], 'ok' if d['self_check']['matching']==d['self_check']['of'] else d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "" "part_bits=10" "part_bits=8" ""; do run c3 "$tune"; done
for tune in "" "part_bits=7" ""; do run c2 "$tune"; done
run c4 ""
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_unpermute -s 3 -c 1 -o gpurun_out/${T}_c3_unperm -f python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_u.log 2>&1; tail -1 gpurun_out/${T}_ncu_u.log | head -c 200; echo

#!/bin/bash
# round 2, call 23: unpartitioned batches through the pipelined in-order kernel (answers written directly)
mkdir -p gpurun_out
T=s23
timeout 900 python -m pytest tests -x -q -m gpu -k "inorder or partitioned_batch or variants or stray or golden or chr3 or query_parity" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
run() {  # workload, tune
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 10 --warmup 3 --cpu-baseline none --e2e-steps 3 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'e2e %.2f' % (d['e2e']['value']/1e9), d['roofline']['kernel']['name'], 'bits', d['roofline']['partition_bits'], 'ok' if d['self_check']['matching']==d['self_check']['of'] else d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "inorder_min=-1" ""; do run c1 "$tune"; done
for tune in "" "part=0" "part=0,inorder_min=-1"; do run c2 "$tune"; done
run small ""

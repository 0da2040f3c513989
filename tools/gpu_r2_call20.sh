#!/bin/bash
# round 2, call 20 (gpurun --gpus 2): k = 31 with the look-ahead first again; replicas and the bit-stream format over 2 GPUs;
# torchrun bench at N = 2
mkdir -p gpurun_out
T=s20
run() {  # workload, tune, extra args
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 $3 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 $3 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'bits', d['roofline']['partition_bits'], d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
run c4 "" "--k 31"
run c3 "" ""
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replicas or bit_stream or partitioned_large" > gpurun_out/${T}_pytest_2gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_2gpu.log
timeout 900 python tools/multi_gpu_api.py > gpurun_out/${T}_multi_api.log 2>&1; tail -45 gpurun_out/${T}_multi_api.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c3_2gpu.log | tail -1 > gpurun_out/${T}_bench_c3_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c3_2gpu.json')); print('torchrun N=2 c3: value %.1f G q/s, e2e %.2f G q/s (bytes %.2f, int64 %.2f)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['byte_api']['value']/1e9, d['e2e']['int64_api']['value']/1e9), d['roofline']['stage_ms'], d['parity'])"

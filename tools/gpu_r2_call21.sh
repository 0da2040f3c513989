#!/bin/bash
# round 2, call 21 (gpurun --gpus 8): the final tree on 8 GPUs: replicas test, N GPUs from one process through the library,
# torchrun bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=s21
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replicas" > gpurun_out/${T}_pytest_replicas.log 2>&1; tail -2 gpurun_out/${T}_pytest_replicas.log
timeout 900 python tools/multi_gpu_api.py > gpurun_out/${T}_multi_api.log 2>&1; grep -E "gpus\"|e2e_Gq|replicate_to|seconds|answers" gpurun_out/${T}_multi_api.log | paste - - - | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c3_${N}gpu.log | tail -1 > gpurun_out/${T}_bench_c3_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c3_${N}gpu.json')); print('torchrun N=$N c3: value %.1f G q/s, e2e %.2f G q/s (bytes %.2f, int64 %.2f)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['byte_api']['value']/1e9, d['e2e']['int64_api']['value']/1e9), d['roofline']['stage_ms'])"

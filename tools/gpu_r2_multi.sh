#!/bin/bash
# multi-GPU call (gpurun --gpus N): replicas inside the library, the host<->device ceiling of the box, torchrun bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=m${N}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replicas" > gpurun_out/${T}_pytest_replicas.log 2>&1; tail -2 gpurun_out/${T}_pytest_replicas.log
timeout 600 python tools/pcie_probe_ranks.py > gpurun_out/${T}_pcie.log 2>&1; tail -40 gpurun_out/${T}_pcie.log | grep -E "gpus|aggregate" | paste - - | head -30
timeout 900 python tools/multi_gpu_api.py > gpurun_out/${T}_multi_api.log 2>&1; tail -45 gpurun_out/${T}_multi_api.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c3.log | tail -1 > gpurun_out/${T}_bench_c3.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c3.json')); print('torchrun N=$N c3: value %.1f G q/s, e2e %.2f G q/s (int64 %.2f)' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['int64_api']['value']/1e9), d['roofline']['stage_ms'])"

#!/bin/bash
# round 2, call 6: full GPU suite, c3 / c2 bench, config 4 (mutated queries; k and nb sweep at 3.1 Gbp), config 5 (align)
mkdir -p gpurun_out
T=s6
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c3.log | tail -1 > gpurun_out/${T}_bench_c3.json
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c2.log | tail -1 > gpurun_out/${T}_bench_c2.json
for cfg in "21 -1" "21 16" "21 20" "21 24" "16 -1" "31 -1" "32 -1"; do
  set -- $cfg
  timeout 900 python bench.py --workload c4 --k $1 --nb $2 --steps 5 --warmup 3 --e2e-steps 2 2> gpurun_out/${T}_bench_c4_k$1_nb$2.log | tail -1 > gpurun_out/${T}_bench_c4_k$1_nb$2.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s6_bench_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f.split("bench_")[1], "%.2f G q/s"%(d["value"]/1e9), "stages", {k: round(v,3) for k,v in r["stage_ms"].items()}, "e2e %.2f"%(d["e2e"]["value"]/1e9), "parity", d["parity"] and {k:v for k,v in d["parity"].items() if "mism" in k or k=="checked" or k=="minus1_answers"}, "self", d["self_check"], "P", round(r["reference_bytes"]["probes_per_query"],2), "nb", d["config"]["nb"], d["config"]["error_bounds"])
    except Exception as e: print(f, "failed", e)
PY
timeout 900 python tools/c5_align_seeds.py > gpurun_out/${T}_c5_seeds.log 2>&1; tail -3 gpurun_out/${T}_c5_seeds.log
timeout 1500 python tools/c5_align_e2e.py > gpurun_out/${T}_c5_e2e.log 2>&1; tail -30 gpurun_out/${T}_c5_e2e.log

#!/bin/bash
# round 2, call 7: in-order kernel variants (what travels a tile ahead x positions kept) x occupancy at c3 and c2;
# the tests that failed in call 6; config 4 at k = 31 / 32; config 5 seeds
mkdir -p gpurun_out
T=s7
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_gpu.log
for occ in 3 4; do for qv in 0 1 2 3 4 5; do
  SAPLING_B200_TUNE="occ=$occ,qv=$qv" timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_c3_occ${occ}_qv${qv}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c3_occ${occ}_qv${qv}.json')); print('c3 occ $occ qv $qv', round(d['roofline']['stage_ms']['query_kernel'],3))"
done; done
for qv in 0 1 2 5; do
  SAPLING_B200_TUNE="occ=4,qv=$qv" timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_c2_occ4_qv${qv}.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c2_occ4_qv${qv}.json')); print('c2 occ 4 qv $qv', round(d['roofline']['stage_ms']['query_kernel'],3))"
done
for cfg in "31 -1" "32 -1"; do
  set -- $cfg
  timeout 900 python bench.py --workload c4 --k $1 --nb $2 --steps 5 --warmup 3 --e2e-steps 2 2> gpurun_out/${T}_bench_c4_k$1_nb$2.log | tail -1 > gpurun_out/${T}_bench_c4_k$1_nb$2.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s7_bench_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f.split("bench_")[1], "%.2f G q/s"%(d["value"]/1e9), "stages", {k: round(v,3) for k,v in r["stage_ms"].items()}, "e2e %.2f"%(d["e2e"]["value"]/1e9), "parity", d["parity"] and {k:v for k,v in d["parity"].items() if "mism" in k or k=="checked" or k=="minus1_answers"}, "self", d["self_check"], "P", round(r["reference_bytes"]["probes_per_query"],2), "nb", d["config"]["nb"], d["config"]["error_bounds"])
    except Exception as e: print(f, "failed", e)
PY
timeout 900 python tools/c5_align_seeds.py > gpurun_out/${T}_c5_seeds.log 2>&1; grep -A8 '"e2e' gpurun_out/${T}_c5_seeds.log | head -30

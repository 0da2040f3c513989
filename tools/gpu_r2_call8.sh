#!/bin/bash
# round 2, call 8: two-sector shortcut in the kernels: tests, c3 / c2 bench, ncu (launch list, full captures)
mkdir -p gpurun_out
T=s8
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c3.log | tail -1 > gpurun_out/${T}_bench_c3.json
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c2.log | tail -1 > gpurun_out/${T}_bench_c2.json
python - <<'PY'
import json
for w in ("c3","c2"):
    try:
        d=json.load(open(f"gpurun_out/s8_bench_{w}.json")); r=d["roofline"]
        print(w, "%.2f G q/s"%(d["value"]/1e9), "stages", {k: round(v,3) for k,v in r["stage_ms"].items()}, "e2e %.2f"%(d["e2e"]["value"]/1e9), "parity", d["parity"] and {k:v for k,v in d["parity"].items() if "mism" in k or k=="checked"}, "frac %.3f"%r["frac"])
    except Exception as e: print(w, "failed", e)
PY
for occ in 3 5; do
  SAPLING_B200_TUNE="occ=$occ" timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/${T}_c3_occ$occ.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c3_occ$occ.json')); print('c3 occ $occ', d['roofline']['stage_ms'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c3_query -f python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c2_query -f python bench.py --workload c2 --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c2.log 2>&1; tail -1 gpurun_out/${T}_ncu_c2.log

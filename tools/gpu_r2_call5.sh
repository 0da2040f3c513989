#!/bin/bash
# round 2, call 5: prediction + L2 prefetch one tile ahead, select-style replay: tests, c3 (occ 5 / 4), c2, c1, ncu
mkdir -p gpurun_out
T=s5
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.log; tail -2 gpurun_out/${T}_bench_c3.log
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.log; tail -2 gpurun_out/${T}_bench_c2.log
timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.log; tail -2 gpurun_out/${T}_bench_c1.log
python - <<'PY'
import json
for w in ("c3","c2","c1"):
    try:
        d=json.load(open(f"gpurun_out/s5_bench_{w}.json"))
        r=d["roofline"]
        print(w, "value %.2f G q/s"%(d["value"]/1e9), "step ms", d["ms_per_step"], "stages", r["stage_ms"], "e2e %.2f / int64 %.2f"%(d["e2e"]["value"]/1e9, d["e2e"]["int64_api"]["value"]/1e9), "parity", d["parity"], "frac", r["frac"], "cpu", d["cpu_baseline"])
    except Exception as e: print(w, "failed", e)
PY
for occ in 4 6; do
  SAPLING_B200_TUNE="occ=$occ" timeout 300 python bench.py --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_c3_occ$occ.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/${T}_c3_occ$occ.json')); print('occ $occ', d['roofline']['stage_ms'])"
done
SAPLING_B200_TUNE="occ=4" timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_c2_occ4.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s5_c2_occ4.json')); print('c2 occ 4', d['roofline']['stage_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c3_query -f python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3.log

#!/bin/bash
# round 2, call 24: the final tree of round 2 (after the in-order direct mode): all GPU tests, smoke, both bench arms at c3, c2 / c1 bench, ncu launch list,
# full captures of the in-order kernel (c3, c2) and of the partition passes
mkdir -p gpurun_out
T=s24
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2> gpurun_out/${T}_bench_c3_ref.log | tail -1 > gpurun_out/${T}_bench_c3_ref.json
( time timeout 900 python bench.py --steps 10 --warmup 3 ) 2> gpurun_out/${T}_bench_c3.log | tail -1 > gpurun_out/${T}_bench_c3.json
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c2.log | tail -1 > gpurun_out/${T}_bench_c2.json
timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c1.log | tail -1 > gpurun_out/${T}_bench_c1.json
timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c4.log | tail -1 > gpurun_out/${T}_bench_c4.json
timeout 900 python bench.py --workload c4 --k 16 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c4_k16.log | tail -1 > gpurun_out/${T}_bench_c4_k16.json
timeout 900 python bench.py --workload c4 --k 31 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_c4_k31.log | tail -1 > gpurun_out/${T}_bench_c4_k31.json
python - <<'PY'
import json
for w in ("c3_ref", "c3", "c2", "c1", "c4", "c4_k16", "c4_k31"):
    try:
        d = json.load(open(f"gpurun_out/s24_bench_{w}.json")); r = d.get("roofline")
        print(w, "%.3f G q/s" % (d["value"] / 1e9), "e2e %.3f" % (d["e2e"]["value"] / 1e9),
              "stages", r and {k: round(v, 3) for k, v in r["stage_ms"].items()}, "frac", r and round(r["frac"], 3),
              "parity", d.get("parity") and {k: v for k, v in d["parity"].items() if "mism" in k or k == "checked"},
              "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(w, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_launches.log 2>&1; tail -1 gpurun_out/${T}_launches.log | head -c 300; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c3_query -f python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c3.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3.log | head -c 300; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmer_query_ordered -s 3 -c 1 -o gpurun_out/${T}_c2_query -f python bench.py --workload c2 --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_c2.log 2>&1; tail -1 gpurun_out/${T}_ncu_c2.log | head -c 300; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_ -s 16 -c 8 -o gpurun_out/${T}_c3_passes -f python bench.py --steps 2 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu_passes.log 2>&1; tail -1 gpurun_out/${T}_ncu_passes.log | head -c 300; echo

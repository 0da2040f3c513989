#!/bin/bash
# round 2, call 28: last check of the final tree: all GPU tests, smoke, both bench arms
mkdir -p gpurun_out
T=s28
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2> gpurun_out/${T}_bench_c3_ref.log | tail -1 > gpurun_out/${T}_bench_c3_ref.json
( time timeout 900 python bench.py ) 2> gpurun_out/${T}_bench_c3.log | tail -1 > gpurun_out/${T}_bench_c3.json
python - <<'PY'
import json
for w in ("c3_ref", "c3"):
    try:
        d = json.load(open(f"gpurun_out/s28_bench_{w}.json")); r = d.get("roofline")
        print(w, "%.3f G q/s" % (d["value"] / 1e9), "steps", d["steps"], "warmup", d["warmup"], "e2e %.3f" % (d["e2e"]["value"] / 1e9),
              "stages", r and {k: round(v, 3) for k, v in r["stage_ms"].items()}, "frac", r and round(r["frac"], 3),
              "parity", d.get("parity"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(w, "failed", e)
PY
tail -4 gpurun_out/${T}_bench_c3.log

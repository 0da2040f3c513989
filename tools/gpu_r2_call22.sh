#!/bin/bash
# round 2, call 22: scatter pass with the write-out as bulk copies (occ=14) against per-thread stores
mkdir -p gpurun_out
T=s22
cat > /tmp/tma_check.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sapling_b200 as S, _fixtures as F
g = F.small_genomes()["rand200k"]
os.environ["SAPLING_B200_TUNE"] = "part=0"
a = S.Sapling.from_memory(g, None, k=21, flags=S.QUIET)
kmers = np.tile(F.query_mix(g, 21, 60000, seed=3), 3)[:150001]
exp = a.queryBatch(kmers); a.close()
for bits in (3, 6, 9, 11):
    os.environ["SAPLING_B200_TUNE"] = f"occ=14,part_min=1,part_bits={bits},chunk_log2=22"
    b = S.Sapling.from_memory(g, None, k=21, flags=S.QUIET)
    got = b.queryBatch(kmers); b.close()
    print("bits", bits, "equal", bool(np.array_equal(got, exp)))
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/tma_check.py 2>&1 | tail -12
SAPLING_B200_TUNE="occ=14" timeout 900 python -m pytest tests -x -q -m gpu -k "partitioned_large or golden or chr3 or bit_stream" > gpurun_out/${T}_pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -2 gpurun_out/${T}_pytest_tma.log
run() {  # workload, tune
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'sustained %.3f ms' % d['sustained']['ms_per_step'], 'bits', d['roofline']['partition_bits'], 'ok' if d['self_check']['matching']==d['self_check']['of'] else d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "occ=14" "" "occ=14" "occ=14,part_bits=10" "occ=14,part_bits=8"; do run c3 "$tune"; done
for tune in "" "occ=14"; do run c2 "$tune"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_scatter -s 3 -c 1 -o gpurun_out/${T}_c3_scatter_tma -f env SAPLING_B200_TUNE="occ=14" python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1; tail -1 gpurun_out/${T}_ncu.log | head -c 200; echo

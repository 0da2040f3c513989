#!/bin/bash
# round 2, call 27: un-permute gather as asynchronous copies into shared memory (u2=1) against the lane-group gather
mkdir -p gpurun_out
T=s27
cat > /tmp/u2_check.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sapling_b200 as S, _fixtures as F
g = F.small_genomes()["rand200k"]
os.environ["SAPLING_B200_TUNE"] = "part=0,inorder_min=-1"
a = S.Sapling.from_memory(g, None, k=21, flags=S.QUIET)
kmers = np.tile(F.query_mix(g, 21, 60000, seed=3), 3)[:150001]
exp = a.queryBatch(kmers); a.close()
for bits in (3, 6, 9, 11):
    os.environ["SAPLING_B200_TUNE"] = f"u2=1,part_min=1,part_bits={bits},chunk_log2=22"
    b = S.Sapling.from_memory(g, None, k=21, flags=S.QUIET)
    got = b.queryBatch(kmers)
    g32 = b.queryBatchU32(kmers); b.close()
    print("bits", bits, "equal", bool(np.array_equal(got, exp)), bool(np.array_equal(np.where(g32 == 0xFFFFFFFF, -1, g32.astype(np.int64)), exp)))
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/u2_check.py 2>&1 | tail -8
timeout 600 compute-sanitizer --tool racecheck python /tmp/u2_check.py 2>&1 | tail -3
SAPLING_B200_TUNE="u2=1" timeout 900 python -m pytest tests -x -q -m gpu -k "partitioned_large or golden or chr3 or bit_stream" > gpurun_out/${T}_pytest_u2.log 2>&1; echo "pytest u2 rc=$?"; tail -2 gpurun_out/${T}_pytest_u2.log
run() {  # workload, tune
  SAPLING_B200_TUNE="$2" timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --cpu-baseline none --e2e-steps 1 2> gpurun_out/${T}_last.log | tail -1 > gpurun_out/${T}_last.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_last.json')); print('$1 [$2]', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, '%.2f G q/s' % (d['value']/1e9), 'sustained %.3f ms' % d['sustained']['ms_per_step'], 'bits', d['roofline']['partition_bits'], 'ok' if d['self_check']['matching']==d['self_check']['of'] else d['self_check'])" || tail -5 gpurun_out/${T}_last.log
}
for tune in "" "u2=1" "" "u2=1" "u2=1,part_bits=10" "u2=1,part_bits=8"; do run c3 "$tune"; done
for tune in "" "u2=1"; do run c2 "$tune"; done

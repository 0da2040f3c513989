#!/usr/bin/env python
"""c3 (3.1 Gbp): inline-prefix suffix array against the plain layout, same index, same queries."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sapling_b200 as S
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 250_000_000
st = torch.cuda.current_stream().cuda_stream
t0 = time.time()
ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10, flags=S.QUIET | S.INLINE)
torch.cuda.synchronize()
print({"build_s": round(time.time() - t0, 2), "device_MB": round(ix.device_bytes() / 1e6), "nb": ix.buckets, "five": ix.five}, flush=True)
d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
rows = []
for mut in (0, 0x5A911C0DE5EED003):
    ix.sample_queries_device(0x5A911C0DE5EED002, mut, 0, nq, d_k.data_ptr(), st)
    ref = None
    for inl, qv in ((0, 3), (1, 3), (1, 4), (1, 5), (1, 6)):
        os.environ["SAPLING_B200_INLINE_QUERY"] = str(inl)
        os.environ["SAPLING_B200_QV"] = str(qv)
        for _ in range(2):
            ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out = d_o.clone()
        if ref is None:
            ref = out
        rows.append({"mutated": bool(mut), "inline": inl, "blocks_per_sm": qv, "ms": round(ms, 2), "Gq_per_s": round(nq / ms / 1e6, 2),
                     "same_results": bool(torch.equal(out, ref))})
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"c3_inline_{n}.json"), "w"), indent=1)

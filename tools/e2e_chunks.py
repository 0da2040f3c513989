#!/usr/bin/env python
"""End-to-end (host buffers) rate of sapling_b200_query_batch against the chunk size of its upload/kernel/download pipeline."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sapling_b200 as S
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10, flags=S.QUIET)
d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), torch.cuda.current_stream().cuda_stream)
h_k = torch.empty(nq, dtype=torch.int64).pin_memory()
h_o = torch.empty(nq, dtype=torch.int64).pin_memory()
h_k.copy_(d_k)
torch.cuda.synchronize()
rows = []
for lg in (0, 19, 20, 21, 22):
    if lg:
        os.environ["SAPLING_B200_CHUNK_LOG2"] = str(lg)
    else:
        os.environ.pop("SAPLING_B200_CHUNK_LOG2", None)
    ix.queryBatch(h_k, out=h_o)
    c0 = ix.launch_count()
    t0 = time.perf_counter()
    for _ in range(5):
        ix.queryBatch(h_k, out=h_o)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    rows.append({"genome_bp": n, "queries": nq, "chunk_log2": lg or "default", "chunks": (ix.launch_count() - c0) // 5,
                 "ms": round(dt * 1e3, 3), "Gq_per_s": round(nq / dt / 1e9, 3), "kernel": ix.query_kernel()[0]})
    print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"e2e_chunks_{n}.json"), "w"), indent=1)

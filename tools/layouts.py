#!/usr/bin/env python
"""Index layouts against each other on the same genome and the same queries (results must be identical):
  plain  : {uint32 suffix array, 2-bit genome}, sector-cached kernel
  inline : 16-byte {position, 27 bases} entries
  packed3: rank lines, overlapping (16 B per base)      packed4: rank lines, tiling (8 B per base)
usage: layouts.py <genome bp> <queries> [layouts, comma separated] [hints, comma separated]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sapling_b200 as S

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
layouts = (sys.argv[3] if len(sys.argv) > 3 else "plain,packed3,packed4").split(",")
hints_list = [int(h) for h in (sys.argv[4] if len(sys.argv) > 4 else "3").split(",")]
st = torch.cuda.current_stream().cuda_stream
d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
rows = []
ref = {}
for lay in layouts:
    for v in ("SAPLING_B200_PACKED_SHIFT", "SAPLING_B200_QV", "SAPLING_B200_HINTS"):
        os.environ.pop(v, None)
    flags = S.QUIET | {"plain": S.NO_PACKED | S.NO_INLINE, "inline": S.INLINE | S.NO_PACKED,
                       "packed3": S.PACKED | S.NO_INLINE, "packed4": S.PACKED | S.NO_INLINE}[lay]
    if lay == "packed4":
        os.environ["SAPLING_B200_PACKED_SHIFT"] = "4"
    for hints in hints_list:
        os.environ["SAPLING_B200_HINTS"] = str(hints)
        t0 = time.time()
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10, flags=flags)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        for mut in (0, 0x5A911C0DE5EED003):
            ix.sample_queries_device(0x5A911C0DE5EED002, mut, 0, nq, d_k.data_ptr(), st)
            variants = [(0, 3), (0, 4), (0, 5)] + ([(1, 2), (1, 3), (1, 4), (1, 5)] if lay.startswith("packed") else [])
            for refill, qv in variants:
                os.environ["SAPLING_B200_QV"] = str(qv)
                os.environ["SAPLING_B200_REFILL"] = str(refill)
                for _ in range(2):
                    ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 4
                if mut not in ref:
                    ref[mut] = d_o.clone()
                rows.append({"genome_bp": n, "queries": nq, "layout": lay, "kernel": ix.query_kernel()[0], "hints": hints,
                             "refill": refill, "mutated_half": bool(mut), "blocks_per_sm": qv, "ms": round(ms, 3),
                             "Gq_per_s": round(nq / ms / 1e6, 2), "same_results": bool(torch.equal(d_o, ref[mut])),
                             "device_MB": round(ix.device_bytes() / 1e6), "build_s": round(build_s, 2)})
                print(rows[-1], flush=True)
            os.environ.pop("SAPLING_B200_QV", None)
            os.environ.pop("SAPLING_B200_REFILL", None)
        ix.close()
        del ix
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"layouts_{n}.json"), "w"), indent=1)

#!/usr/bin/env python
"""GPU experiments behind the design decisions in DESIGN.md (run under gpurun; prints a small report).

  python tools/gpu_experiments.py gather     # random-access microbenchmarks (roofline denominators)
  python tools/gpu_experiments.py variants   # query-kernel occupancy variants on the c2 workload
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sapling_b200 as S  # noqa: E402


def gather():
    rows = []
    for foot_gb in (0.4, 12):
        nbytes = int(foot_gb * (1 << 30))
        for gran in (32, 64, 128):
            for chain in (1, 5):
                for bps in (8,):
                    g = S.gather_bench2(nbytes, 1 << 28, gran, chain, bps, reps=2)
                    rows.append({"footprint_GB": foot_gb, "gran_B": gran, "chain": chain, "blocks_per_sm": bps,
                                 "Gaccess_per_s": round(g, 2), "GB_per_s": round(g * gran, 1)})
                    print(rows[-1], flush=True)
    for bps in (4, 6):
        g = S.gather_bench2(12 << 30, 1 << 28, 32, 5, bps, reps=2)
        rows.append({"footprint_GB": 12, "gran_B": 32, "chain": 5, "blocks_per_sm": bps,
                     "Gaccess_per_s": round(g, 2), "GB_per_s": round(g * 32, 1)})
        print(rows[-1], flush=True)
    return rows


def _time_queries(ix, d_k, d_o, nq, st, reps=5):
    import torch
    for _ in range(3):
        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def variants(workload_n=100_000_000, nq=50_000_000):
    """Layout / L2-hint / pipelining / occupancy sweep of the query kernel on the c2 workload."""
    import torch
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows, ref = [], None
    index_cfgs = [(0, 0), (1, 0), (1, 15)]          # (narrow, hints): fixed at index creation
    launch_cfgs = [(0, 4, 1), (0, 4, 2), (1, 3, 1), (1, 4, 1), (1, 4, 2), (1, 5, 1), (1, 3, 2), (1, 6, 1)]  # (pipeline, qv, mult)
    for narrow, hints in index_cfgs:
        os.environ["SAPLING_B200_NARROW"] = str(narrow)
        os.environ["SAPLING_B200_HINTS"] = str(hints)
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, workload_n, k=21, maxMem=10)
        ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
        torch.cuda.synchronize()
        for pipe, qv, mult in launch_cfgs:
            if pipe and not narrow:
                continue
            os.environ["SAPLING_B200_PIPELINE"] = str(pipe)
            os.environ["SAPLING_B200_QV"] = str(qv)
            os.environ["SAPLING_B200_GRID_MULT"] = str(mult)
            ms = _time_queries(ix, d_k, d_o, nq, st)
            out = d_o.cpu()
            if ref is None:
                ref = out
            rows.append({"narrow": narrow, "hints": hints, "pipeline": pipe, "blocks_per_sm": qv, "grid_mult": mult,
                         "ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 2), "same_results": bool(torch.equal(out, ref))})
            print(rows[-1], flush=True)
        ix.close()
    return rows


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gather"
    t0 = time.time()
    res = {"gather": gather, "variants": variants}[what]()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"exp_{what}.json"), "w"), indent=1)
    print(f"done in {time.time() - t0:.1f}s")

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


def footprint():
    """Random-access throughput against footprint: locates the L2-capacity and TLB-reach knees."""
    rows = []
    for mb in (16, 48, 96, 192, 400, 800, 1600, 3200, 12800, 51200):
        for gran in (16, 64):
            g = S.gather_bench2(mb << 20, 1 << 28, gran, 1, 8, reps=2)
            rows.append({"footprint_MB": mb, "gran_B": gran, "Gaccess_per_s": round(g, 2), "GB_per_s": round(g * gran, 1)})
            print(rows[-1], flush=True)
    return rows


def tlb():
    """Is DRAM-resident random access limited by address translation?  Same 12.8 GB / 51 GB buffers, but each block
    confined to one of `slices` contiguous slices (a few pages per SM) -- DRAM rows and L2 stay just as cold."""
    rows = []
    for mb in (400, 12800, 51200):
        for slices in (1, 8, 148, 1184, 148 * 64):
            os.environ["SAPLING_B200_GATHER_SLICES"] = str(slices)
            for gran in (16, 64):
                g = S.gather_bench2(mb << 20, 1 << 28, gran, 1, 8, reps=2)
                rows.append({"footprint_MB": mb, "slices": slices, "gran_B": gran, "Gaccess_per_s": round(g, 2),
                             "GB_per_s": round(g * gran, 1)})
                print(rows[-1], flush=True)
    os.environ.pop("SAPLING_B200_GATHER_SLICES", None)
    return rows


def hints():
    """Which arrays to pin in L2 (the effective capacity for data shared by all SMs is ~60 MB, see footprint())."""
    import torch
    nq = 50_000_000
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for n in (100_000_000,):
        for h in (13, 5, 1, 9, 12, 4, 3, 15, 0, 7):
            os.environ["SAPLING_B200_HINTS"] = str(h)
            ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10)
            ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
            torch.cuda.synchronize()
            for ln, pipe, qv, mult in ((0, 0, 4, 2), (0, 1, 3, 2), (1, 0, 4, 2), (1, 0, 5, 2)):
                os.environ.update({"SAPLING_B200_LINE": str(ln), "SAPLING_B200_PIPELINE": str(pipe),
                                   "SAPLING_B200_QV": str(qv), "SAPLING_B200_GRID_MULT": str(mult)})
                ms = _time_queries(ix, d_k, d_o, nq, st)
                rows.append({"n": n, "hints": h, "line": ln, "pipeline": pipe, "blocks_per_sm": qv, "grid_mult": mult,
                             "ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 2)})
                print(rows[-1], flush=True)
            ix.close()
    return rows


def sizes():
    """Throughput against genome size (index footprint): where the L2-capacity and TLB-reach knees are."""
    import torch
    nq = 50_000_000
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    os.environ["SAPLING_B200_HINTS"] = "15"
    for n in (5_000_000, 10_000_000, 25_000_000, 50_000_000, 100_000_000, 200_000_000, 400_000_000, 800_000_000):
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10)
        ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
        torch.cuda.synchronize()
        for ln, pipe, qv, mult in ((0, 0, 4, 2), (1, 0, 4, 2)):
            os.environ.update({"SAPLING_B200_LINE": str(ln), "SAPLING_B200_PIPELINE": str(pipe),
                               "SAPLING_B200_QV": str(qv), "SAPLING_B200_GRID_MULT": str(mult)})
            ms = _time_queries(ix, d_k, d_o, nq, st)
            rows.append({"n": n, "nb": ix.buckets, "five": list(ix.five), "device_MB": round(ix.device_bytes() / 1e6),
                         "line": ln, "ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 2)})
            print(rows[-1], flush=True)
        ix.close()
    return rows


def pin():
    """L2 pinning for the default (sector) kernel on c2: eviction hints x persisting-L2 carve-out x occupancy."""
    import torch
    nq = 50_000_000
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for persist in (0, 40, 79):
        for h in (3, 4, 5, 7, 13, 15):
            os.environ["SAPLING_B200_HINTS"] = str(h)
            os.environ["SAPLING_B200_L2_PERSIST_MB"] = str(persist)
            ix = S.Sapling.synthetic(0x5A911C0DE5EED001, 100_000_000, k=21, maxMem=10)
            ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
            torch.cuda.synchronize()
            for qv in (4, 5):
                os.environ.update({"SAPLING_B200_QV": str(qv), "SAPLING_B200_GRID_MULT": "2"})
                ms = _time_queries(ix, d_k, d_o, nq, st)
                rows.append({"persist_MB": persist, "hints": h, "blocks_per_sm": qv, "ms": round(ms, 3),
                             "Gq_per_s": round(nq / ms / 1e6, 2)})
                print(rows[-1], flush=True)
            ix.close()
    os.environ["SAPLING_B200_L2_PERSIST_MB"] = "0"
    return rows


def nbsweep():
    """c2 genome with other bucket counts: how much a smaller model table (L2 residency) buys, probes held ~constant."""
    import torch
    nq = 50_000_000
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for persist, h in ((0, 3), (40, 1), (79, 3)):
        os.environ["SAPLING_B200_HINTS"] = str(h)
        os.environ["SAPLING_B200_L2_PERSIST_MB"] = str(persist)
        for nb in (19, 20, 21, 22, 23, 24, 25):
            ix = S.Sapling.synthetic(0x5A911C0DE5EED001, 100_000_000, numBuckets=nb, k=21, maxMem=10)
            ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
            torch.cuda.synchronize()
            for qv in (4, 5):
                os.environ.update({"SAPLING_B200_QV": str(qv), "SAPLING_B200_GRID_MULT": "2"})
                ms = _time_queries(ix, d_k, d_o, nq, st)
                rows.append({"persist_MB": persist, "hints": h, "nb": nb, "five": list(ix.five), "model_MB": (8 << nb) >> 20,
                             "blocks_per_sm": qv, "ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 2)})
                print(rows[-1], flush=True)
            ix.close()
    os.environ["SAPLING_B200_L2_PERSIST_MB"] = "0"
    return rows


def kernels():
    """Kernel variants on c2 and c3-like sizes with the default index configuration."""
    import torch
    rows = []
    st = torch.cuda.current_stream().cuda_stream
    for n, nq in ((100_000_000, 50_000_000), (10_000_000, 50_000_000), (3_100_000_000, 250_000_000)):
        d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
        d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, n, k=21, maxMem=10)
        ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
        torch.cuda.synchronize()
        ref = None
        for sector, line, pipe, qv, mult in ((0, 0, 0, 4, 2), (1, 0, 0, 4, 2), (1, 0, 0, 4, 1), (1, 0, 0, 3, 2), (1, 0, 0, 5, 2),
                                             (1, 0, 0, 4, 4)):
            os.environ.update({"SAPLING_B200_SECTOR": str(sector), "SAPLING_B200_LINE": str(line),
                               "SAPLING_B200_PIPELINE": str(pipe), "SAPLING_B200_QV": str(qv),
                               "SAPLING_B200_GRID_MULT": str(mult)})
            ms = _time_queries(ix, d_k, d_o, nq, st, reps=3)
            out = d_o.cpu()
            if ref is None:
                ref = out
            rows.append({"n": n, "sector": sector, "line": line, "pipeline": pipe, "blocks_per_sm": qv, "grid_mult": mult,
                         "ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 2), "same_results": bool(torch.equal(out, ref))})
            print(rows[-1], flush=True)
        ix.close()
        del d_k, d_o
        torch.cuda.empty_cache()
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
    """Layout / L2-hint / line-cache / pipelining / occupancy sweep of the query kernel on the c2 workload."""
    import torch
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    p = torch.cuda.get_device_properties(0)
    print({"l2_bytes": p.L2_cache_size, "sms": p.multi_processor_count}, flush=True)
    rows, ref = [], None
    # (narrow, hints, fetch granularity, persisting-L2 MB): fixed at index creation
    index_cfgs = [(1, 15, None, None), (1, 0, None, None), (1, 3, None, None), (1, 11, None, None),
                  (1, 15, 32, None), (1, 15, 128, None), (1, 15, 64, 64), (0, 15, 64, 0)]
    launch_cfgs = [(ln, pipe, qv, mult) for ln in (1, 0) for pipe in (1, 0) for qv in (3, 4, 5) for mult in (1, 2)]
    for narrow, hints, fetch, persist in index_cfgs:
        os.environ["SAPLING_B200_NARROW"] = str(narrow)
        os.environ["SAPLING_B200_HINTS"] = str(hints)
        for name, v in (("SAPLING_B200_L2_FETCH", fetch), ("SAPLING_B200_L2_PERSIST_MB", persist)):
            if v is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = str(v)
        ix = S.Sapling.synthetic(0x5A911C0DE5EED001, workload_n, k=21, maxMem=10)
        ix.sample_queries_device(0x5A911C0DE5EED002, 0, 0, nq, d_k.data_ptr(), st)
        torch.cuda.synchronize()
        for ln, pipe, qv, mult in launch_cfgs:
            if (pipe and not narrow) or (narrow and (fetch or persist)) and not (ln and qv == 4 and mult == 1):
                continue
            os.environ["SAPLING_B200_LINE"] = str(ln)
            os.environ["SAPLING_B200_PIPELINE"] = str(pipe)
            os.environ["SAPLING_B200_QV"] = str(qv)
            os.environ["SAPLING_B200_GRID_MULT"] = str(mult)
            ms = _time_queries(ix, d_k, d_o, nq, st)
            out = d_o.cpu()
            if ref is None:
                ref = out
            rows.append({"narrow": narrow, "hints": hints, "fetch": fetch, "persist": persist, "line": ln,
                         "pipeline": pipe, "blocks_per_sm": qv, "grid_mult": mult, "ms": round(ms, 3),
                         "Gq_per_s": round(nq / ms / 1e6, 2), "same_results": bool(torch.equal(out, ref))})
            print(rows[-1], flush=True)
        ix.close()
    # restore defaults for whatever runs next in this process
    for name in ("SAPLING_B200_L2_FETCH", "SAPLING_B200_L2_PERSIST_MB", "SAPLING_B200_LINE", "SAPLING_B200_PIPELINE",
                 "SAPLING_B200_QV", "SAPLING_B200_GRID_MULT", "SAPLING_B200_NARROW", "SAPLING_B200_HINTS"):
        os.environ.pop(name, None)
    return rows


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gather"
    t0 = time.time()
    res = {"gather": gather, "variants": variants, "footprint": footprint, "tlb": tlb, "hints": hints, "sizes": sizes, "kernels": kernels, "pin": pin, "nbsweep": nbsweep}[what]()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"exp_{what}.json"), "w"), indent=1)
    print(f"done in {time.time() - t0:.1f}s")

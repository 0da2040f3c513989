#!/usr/bin/env python
"""First contact with the c3 configuration (BASELINE.json configs[2]): synthetic 3.1 Gbp genome, suffix array and
model built on the GPU, k=21 queries; compat (the reference's (int)predicted arithmetic) against the
64-bit-safe mode; parity of a sample against the oracle port built from the same parts.

  python tools/c3_probe.py [n] [nq] [parity_sample]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sapling_b200 as S  # noqa: E402

SEED_G, SEED_Q, SEED_M = 0x5A911C0DE5EED001, 0x5A911C0DE5EED002, 0x5A911C0DE5EED003


def timeit(ix, d_k, d_o, nq, st, reps=3):
    ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.queryBatchDevice(d_k.data_ptr(), nq, d_o.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
    nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 250_000_000
    ps = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2_000_000
    res = {"n": n, "nq": nq}
    st = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    ix = S.Sapling.synthetic(SEED_G, n, k=21, maxMem=10, keep_host_genome=ps > 0, flags=S.QUIET)
    torch.cuda.synchronize()
    res["build_s"] = round(time.time() - t0, 2)
    res.update({"nb": ix.buckets, "five": list(ix.five), "device_MB": round(ix.device_bytes() / 1e6),
                "peak_alloc_GB": None})
    print(res, flush=True)
    d_k = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_o = torch.empty(nq, dtype=torch.int64, device="cuda")
    for name, mut in (("present", 0), ("half_mutated", SEED_M)):
        ix.sample_queries_device(SEED_Q, mut, 0, nq, d_k.data_ptr(), st)
        torch.cuda.synchronize()
        ms = timeit(ix, d_k, d_o, nq, st)
        match, m1 = ix.verify_device(d_k.data_ptr(), d_o.data_ptr(), nq, st)
        res[name] = {"ms": round(ms, 2), "Gq_per_s": round(nq / ms / 1e6, 2), "matching": match, "minus1": m1}
        print(name, res[name], flush=True)
        if ps > 0:
            import _oracle as O
            if "port" not in res:
                t0 = time.time()
                xl, yl = ix.model()
                port = O.Port.from_parts(ix.reference, ix.rev(), 21, ix.buckets, xl, yl, ix.five)
                res["port"] = f"oracle port from parts in {time.time() - t0:.0f}s"
                print(res["port"], flush=True)
            # sample spread over the whole batch: high ranks (>= 2^31) matter (F5)
            samp = d_k[:: max(1, nq // ps)][:ps].contiguous()
            got = torch.empty(len(samp), dtype=torch.int64, device="cuda")
            ix.queryBatchDevice(samp.data_ptr(), len(samp), got.data_ptr(), st)
            torch.cuda.synchronize()
            t0 = time.time()
            exp, probes, oob = port.query_batch(samp.cpu().numpy().astype(np.uint64), nthreads=os.cpu_count(), stats=True)
            dt = time.time() - t0
            g = got.cpu().numpy()
            res[name]["parity"] = {"checked": len(samp), "mismatches": int((g != exp).sum()),
                                   "probes_per_query": round(probes / len(samp), 3), "oob": int(oob),
                                   "oracle_Mq_per_s": round(len(samp) / dt / 1e6, 2), "threads": os.cpu_count()}
            print(name, "parity", res[name]["parity"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"c3_probe_{n}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Unpartitioned batches: the pipelined in-order kernel against one query per thread, by batch size (10 Mbp index, device-
resident queries, CUDA events, median of 30).  Decides Tuning::inorder_min (capi.cu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import sapling_b200 as S
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    rows = []
    idx = {}
    for name, tune in (("inorder", "part=0,inorder_min=1"), ("per_thread", "part=0,inorder_min=-1")):
        os.environ["SAPLING_B200_TUNE"] = tune
        idx[name] = S.Sapling.synthetic(0x5EED0001C0FFEE01, n, k=21, maxMem=10, flags=S.QUIET)
    nmax = 1 << 22
    d = torch.empty(nmax, dtype=torch.int64, device="cuda")
    out = torch.empty(nmax, dtype=torch.int64, device="cuda")
    idx["inorder"].sample_queries_device(0x5EED0002BADC0DE5, 0, 0, nmax, d.data_ptr(), 0)
    torch.cuda.synchronize()
    for lg in range(10, 23):
        nq = 1 << lg
        row = {"nq": nq}
        for name, ix in idx.items():
            ts = []
            for _ in range(35):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ix.queryBatchDevice(d.data_ptr(), nq, out.data_ptr(), 0)
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts = sorted(ts[5:])
            row[name + "_us"] = round(1e3 * ts[len(ts) // 2], 1)
        rows.append(row)
        print(row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n": n, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "inorder_crossover.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""The host<->device ceiling of this box when 1 .. N GPUs copy at once (pinned memory, both directions together): what bounds
the end-to-end path at N GPUs.  One thread per GPU in one process (as sapling_b200_query_batch does).

  python tools/pcie_probe_ranks.py
"""
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = 1_000_000_000
ngpu = torch.cuda.device_count()
bufs = []
for g in range(ngpu):
    with torch.cuda.device(g):
        bufs.append((torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                     torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"), torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"),
                     torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))


def work(g, reps, mode):
    h_in, h_out, d_a, d_b, s1, s2 = bufs[g]
    with torch.cuda.device(g):
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        s1.synchronize()
        s2.synchronize()


res = {"gpus_visible": ngpu, "bytes_per_copy": n, "rows": []}
for g in [x for x in (1, 2, 4, 8) if x <= ngpu]:
    for mode in ("h2d", "d2h", "both"):
        for reps in (1, 4):  # warm-up, then timed
            th = [threading.Thread(target=work, args=(i, reps, mode)) for i in range(g)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
        moved = g * reps * n * (2 if mode == "both" else 1)
        res["rows"].append({"gpus": g, "mode": mode, "aggregate_GB_per_s": round(moved / dt / 1e9, 1),
                            "per_gpu_GB_per_s": round(moved / dt / 1e9 / g, 1)})
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"pcie_probe_{ngpu}.json"), "w"), indent=1)

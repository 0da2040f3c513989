#!/usr/bin/env python
"""Multi-GPU INSIDE the library: one process, one handle.  The index is built once on GPU 0, copied device to
device onto the other GPUs (sapling_b200_replicate: cudaMemcpyPeer over NVLink / NVSwitch), and the host-pointer batch call
shards every batch over all of them with one host thread per GPU.  Reports the replication time / rate and the end-to-end
rate of sapling_b200_query_batch_u32 at 1 .. N GPUs, with the answers of every configuration compared.

  python tools/multi_gpu_api.py [n=3.1e9] [nq=2.5e8]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sapling_b200 as S  # noqa: E402

SEED_G, SEED_Q, K = 0x5A911C0DE5EED001, 0x5A911C0DE5EED002, 21


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
    nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 250_000_000
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(0)
    t0 = time.time()
    ix = S.Sapling.synthetic(SEED_G, n, k=K, maxMem=10, flags=S.QUIET)
    res = {"n": n, "nq": nq, "gpus_visible": ngpu, "build_s": round(time.time() - t0, 2),
           "index_bytes_per_gpu": ix.device_bytes(), "runs": []}
    d = torch.empty(nq, dtype=torch.int64, device="cuda")
    ix.sample_queries_device(SEED_Q, 0, 0, nq, d.data_ptr(), 0)
    torch.cuda.synchronize()
    sys.path.insert(0, ROOT)
    from bench import pack_kmer_bits_device
    bits = 2 * K
    packed = torch.empty((nq * bits + 7) // 8, dtype=torch.uint8).pin_memory()
    packed.copy_(pack_kmer_bits_device(d, bits))  # the densest upload format: nothing but the k-mers
    torch.cuda.synchronize()
    out = torch.empty(nq, dtype=torch.int32).pin_memory()
    del d
    first = None
    have = 1
    for g in [x for x in (1, 2, 4, 8) if x <= ngpu]:
        if g > have:
            t0 = time.time()
            ix.replicate((1 << g) - 1)
            dt = time.time() - t0
            res["runs"].append({"replicate_to_gpus": g, "seconds": round(dt, 2),
                                "GB_per_s_per_new_gpu": round(ix.device_bytes() / dt / 1e9, 1)})
            have = g
        for _ in range(2):
            ix.queryBatchBits(packed, bits, nq, out=out)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ix.queryBatchBits(packed, bits, nq, out=out)
        dt = (time.perf_counter() - t0) / reps
        cur = out.clone()
        if first is None:
            first = cur
        res["runs"].append({"gpus": g, "e2e_Gq_per_s": round(nq / dt / 1e9, 2), "ms": round(dt * 1e3, 1),
                            "answers_equal_1gpu": bool(torch.equal(cur, first)), "bytes_per_query": bits / 8 + 4})
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_api_{ngpu}.json"), "w"), indent=1)
    ix.close()


if __name__ == "__main__":
    main()

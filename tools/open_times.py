#!/usr/bin/env python
"""How long an index takes to open, by source, at the largest genome the box's scratch disk holds (3.1 Gbp needs ~75 GB):

  built on the GPU from the bases alone        Sapling.synthetic
  the reference's files (FASTA + .sa + .sap)   Sapling(refFn, saFn, sapFn)   -- what the reference constructor reads
  the private cache next to them               Sapling.from_cache            -- SAPLING_B200_CACHE=1 / sapling_b200_open_cache

and, with more than one GPU, a replica (device to device) against opening the files once more.  Writes
gpurun_out/open_times.json.  Nothing here is on the product path."""
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SEED_G = 0x5EED0001C0FFEE01


def main():
    import torch
    import sapling_b200 as S
    scratch = sys.argv[1] if len(sys.argv) > 1 else "/tmp/sapling_open"
    os.makedirs(scratch, exist_ok=True)
    free = shutil.disk_usage(scratch).free
    n = 3_100_000_000
    while n * 25 > free * 0.9 and n > 10_000_000:   # FASTA 1 + .sa 16 + .sap ~1.4 + cache 5.6 bytes per base
        n //= 2
    # and no larger than the disk writes in about two minutes (probe: 1 GB)
    t0 = time.time()
    with open(os.path.join(scratch, "probe"), "wb") as f:
        blk = np.random.default_rng(1).integers(0, 255, size=1 << 26, dtype=np.uint8).tobytes()
        for _ in range(16):
            f.write(blk)
        f.flush()
        os.fsync(f.fileno())
    write_gbs = (16 * (1 << 26)) / (time.time() - t0) / 1e9
    os.remove(os.path.join(scratch, "probe"))
    while n * 25 / (write_gbs * 1e9) > 130 and n > 10_000_000:
        n //= 2
    res = {"scratch": scratch, "disk_free_GB": round(free / 1e9, 1), "disk_write_GB_per_s": round(write_gbs, 2), "n": n, "k": 21,
           "gpus": torch.cuda.device_count()}
    t0 = time.time()
    ix = S.Sapling.synthetic(SEED_G, n, k=21, maxMem=10, keep_host_genome=True, flags=S.QUIET | S.KEEP_BUILD)
    res["build_on_gpu_s"] = round(time.time() - t0, 2)
    res["resident_GB"] = round(ix.device_bytes() / 1e9, 2)
    fa, sa, sap = (os.path.join(scratch, x) for x in ("g.fa", "g.fa.sa", "g.fa.sap"))
    g = np.frombuffer(ix.reference, dtype=np.uint8)
    t0 = time.time()
    with open(fa, "wb") as f:
        f.write(b">chr1\n")
        W = 80
        full = (len(g) // W) * W
        step = W * (1 << 22)
        for o in range(0, full, step):
            blk = g[o:min(o + step, full)].reshape(-1, W)
            f.write(np.concatenate([blk, np.full((blk.shape[0], 1), 10, dtype=np.uint8)], axis=1).tobytes())
        if full < len(g):
            f.write(g[full:].tobytes() + b"\n")
    res["write_fasta_s"] = round(time.time() - t0, 2)
    t0 = time.time(); ix.write_sa(sa); res["write_sa_s"] = round(time.time() - t0, 2)
    t0 = time.time(); ix.write_sap(sap); res["write_sap_s"] = round(time.time() - t0, 2)
    t0 = time.time(); ix.save_cache(sap + ".b200"); res["save_cache_s"] = round(time.time() - t0, 2)
    res["bytes"] = {os.path.basename(p): os.path.getsize(p) for p in (fa, sa, sap, sap + ".b200")}
    kmers = torch.empty(1 << 20, dtype=torch.int64, device="cuda")
    ix.sample_queries_device(0x5EED0002BADC0DE5, 0, 0, kmers.numel(), kmers.data_ptr(), 0)
    torch.cuda.synchronize()
    hk = kmers.cpu().numpy().view(np.uint64)
    want = ix.queryBatch(hk)
    five, nb = ix.five, ix.buckets
    ix.close()
    del g
    os.sync()
    for label, opener in (("open_reference_files_s", lambda: S.Sapling(fa, sa, sap, k=21, flags=S.QUIET)),
                          ("open_reference_files_again_s", lambda: S.Sapling(fa, sa, sap, k=21, flags=S.QUIET)),
                          ("open_cache_s", lambda: S.Sapling.from_cache(sap + ".b200", flags=S.QUIET))):
        t0 = time.time()
        jx = opener()
        res[label] = round(time.time() - t0, 2)
        assert jx.five == five and jx.buckets == nb and np.array_equal(jx.queryBatch(hk), want), label
        if label == "open_cache_s" and res["gpus"] > 1:
            t0 = time.time()
            jx.replicate((1 << res["gpus"]) - 1)
            res["replicate_to_all_gpus_s"] = round(time.time() - t0, 2)
        jx.close()
    shutil.rmtree(scratch, ignore_errors=True)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "open_times.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json configs[4]: align seed lookup.  1 M synthetic 150 bp reads (1 % substitutions, a few N, a third
reverse-complemented), num_seeds=7, sapling_k=16, max_hits=32 -> 14 M seed queries per block through
sapling_b200_seed_batch; parity of a read sample against the oracle (and the reference's own methods when built).

  python tools/c5_align_seeds.py [n=1e8] [n_reads=1e6] [parity_reads=5e4]
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
import _oracle as O  # noqa: E402

SEED_G, SEED_R = 0x5A911C0DE5EED001, 0x5A911C0DE5EED005
K, NUM_SEEDS, MAX_HITS, LEN = 16, 7, 32, 150


def make_reads(genome: bytes, n_reads):
    """Vectorised read simulator: start = splitmix64(seed+j) mod (n-LEN); 1 % substitutions; 0.05 % N; every third read
    reverse-complemented."""
    g = np.frombuffer(genome, dtype=np.uint8)
    rng = np.random.default_rng(12345)
    with np.errstate(over="ignore"):
        starts = (O.splitmix64_np(np.uint64(SEED_R) + np.arange(n_reads, dtype=np.uint64)) % np.uint64(len(g) - LEN)).astype(np.int64)
    reads = g[starts[:, None] + np.arange(LEN)[None, :]]
    m = rng.random(reads.shape) < 0.01
    reads[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
    reads[rng.random(reads.shape) < 0.0005] = ord("N")
    comp = np.arange(256, dtype=np.uint8)
    for a, b in (b"AT", b"CG", b"GC", b"TA"):
        comp[a] = b
    rc = np.arange(n_reads) % 3 == 2
    reads[rc] = comp[reads[rc][:, ::-1]]
    return np.ascontiguousarray(reads), starts


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
    ps = int(float(sys.argv[3])) if len(sys.argv) > 3 else 50_000
    t0 = time.time()
    ix = S.Sapling.synthetic(SEED_G, n, k=K, maxMem=10, keep_host_genome=True, flags=S.QUIET | S.KEEP_BUILD)
    genome = ix.reference
    res = {"n": n, "k": K, "nb": ix.buckets, "five": list(ix.five), "n_reads": n_reads, "read_len": LEN,
           "num_seeds": NUM_SEEDS, "max_hits": MAX_HITS, "seeds_per_block": n_reads * 2 * NUM_SEEDS,
           "index_build_s": round(time.time() - t0, 2)}
    reads, starts = make_reads(genome, n_reads)
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(LEN))
    total = n_reads * 2 * NUM_SEEDS
    # device-resident
    st = torch.cuda.current_stream().cuda_stream
    d_reads = torch.from_numpy(reads.reshape(-1)).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_rp = torch.empty(total, dtype=torch.int32, device="cuda")   # compact tuples: 4 + 4 + 1 + 1 bytes per seed
    d_sp = torch.empty(total, dtype=torch.int32, device="cuda")
    d_l, d_r = (torch.empty(total, dtype=torch.uint8, device="cuda") for _ in range(2))
    args = (d_reads.data_ptr(), d_off.data_ptr(), n_reads, NUM_SEEDS, MAX_HITS, d_rp.data_ptr(), d_sp.data_ptr(),
            d_l.data_ptr(), d_r.data_ptr(), st)
    for _ in range(3):
        ix.seedBatchDevice(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ix.seedBatchDevice(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["device"] = {"ms": round(ms, 3), "Gseeds_per_s": round(total / ms / 1e6, 2), "Mreads_per_s": round(n_reads / ms / 1e3, 1)}
    # end to end through the host C ABI (reads on the host, tuples back on the host)
    read_list = None
    blob = reads.tobytes()
    rp32, sp = np.empty(total, np.uint32), np.empty(total, np.uint32)
    lf8, rt8 = np.empty(total, np.uint8), np.empty(total, np.uint8)
    L = S.lib()
    L.sapling_b200_seed_batch_compact(ix._h, blob, off, n_reads, NUM_SEEDS, MAX_HITS, rp32, sp, lf8, rt8)
    t0 = time.perf_counter()
    for _ in range(3):
        assert L.sapling_b200_seed_batch_compact(ix._h, blob, off, n_reads, NUM_SEEDS, MAX_HITS, rp32, sp, lf8, rt8) == 0
    dt = (time.perf_counter() - t0) / 3
    # the same call with pinned caller buffers (copied from / to directly)
    p_blob = torch.from_numpy(reads.reshape(-1).copy()).pin_memory()
    p_rp, p_sp = torch.empty(total, dtype=torch.int32).pin_memory(), torch.empty(total, dtype=torch.int32).pin_memory()
    p_l, p_r = torch.empty(total, dtype=torch.uint8).pin_memory(), torch.empty(total, dtype=torch.uint8).pin_memory()
    import ctypes as C
    raw = C.CDLL(S.lib_path())
    raw.sapling_b200_seed_batch_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    call = lambda: raw.sapling_b200_seed_batch_compact(ix._h, p_blob.data_ptr(), off.ctypes.data, n_reads, NUM_SEEDS, MAX_HITS,
                                                       p_rp.data_ptr(), p_sp.data_ptr(), p_l.data_ptr(), p_r.data_ptr())
    assert call() == 0
    t0 = time.perf_counter()
    for _ in range(3):
        assert call() == 0
    dtp = (time.perf_counter() - t0) / 3
    assert np.array_equal(p_rp.numpy().view(np.uint32), rp32)
    res["e2e_pinned"] = {"ms": round(dtp * 1e3, 2), "Gseeds_per_s": round(total / dtp / 1e9, 3),
                         "Mreads_per_s": round(n_reads / dtp / 1e6, 2)}
    res["e2e"] = {"ms": round(dt * 1e3, 2), "Gseeds_per_s": round(total / dt / 1e9, 3), "Mreads_per_s": round(n_reads / dt / 1e6, 2),
                  "h2d_bytes": len(blob) + off.nbytes, "d2h_bytes": total * 10,
                  "api": "sapling_b200_seed_batch_compact (blocks of reads pipelined, 10 bytes per seed back)"}
    assert np.array_equal(rp32, d_rp.cpu().numpy().view(np.uint32))
    rp = np.where(rp32 == 0xFFFFFFFF, -1, rp32.astype(np.int64))
    lf, rt = lf8.astype(np.uint32), rt8.astype(np.uint32)
    hits = rp.reshape(n_reads, 2, NUM_SEEDS) >= 0
    res["reads_with_a_hit"] = int(hits.any(axis=(1, 2)).sum())
    res["seed_hits"] = int(hits.sum())
    # parity + CPU arm on a read sample
    sample = [reads[i].tobytes() for i in range(ps)]
    xl, yl = ix.model()
    threads = os.cpu_count()
    port = O.Port.from_memory(genome, sa=ix.rev(), nb=ix.buckets, k=K)      # needs inv + lcp for the hit counts
    assert port.five == ix.five
    t0 = time.perf_counter()
    exp = port.seed_batch(sample, NUM_SEEDS, MAX_HITS, nthreads=threads)
    dt = time.perf_counter() - t0
    m = ps * 2 * NUM_SEEDS
    got = (rp[:m].reshape(ps, 2, NUM_SEEDS), sp[:m].reshape(ps, 2, NUM_SEEDS), lf[:m].reshape(ps, 2, NUM_SEEDS),
           rt[:m].reshape(ps, 2, NUM_SEEDS))
    res["parity"] = {"reads": ps, "seeds": m, "mismatches": {w: int((a != b).sum()) for a, b, w in
                                                             zip(got, exp, ("ref_pos", "sa_pos", "left", "right"))}}
    res["cpu_baseline"] = {"kind": "port", "cores": threads, "Mseeds_per_s": round(m / dt / 1e6, 2),
                           "sample": f"{ps} reads, OpenMP over the oracle's align.cpp:259-300 restatement"}
    port.close()
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "c5_align_seeds.json"), "w"), indent=1)
    ix.close()


if __name__ == "__main__":
    main()

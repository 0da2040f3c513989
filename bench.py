#!/usr/bin/env python
"""bench.py -- k=21 SAPLING suffix-array queries/sec on B200.

Default workload (BASELINE.json configs[2], "c3", the configuration the metric is quoted on): synthetic 3.1 Gbp
random-ACGT genome (counter-based generator), suffix array and .sap model built on the GPU with the
reference's default parameters (k=21, maxMem=10 -> nb=28), 250 M 21-mers sampled from the genome per GPU per step
(4 GPU-steps = the 1 B queries of the config).  c2 (configs[1], 100 Mbp / 50 M), c1, c4 (configs[3]: 50 % of the queries
carry 1-2 substitutions, k and nb selectable) and `small` stay selectable.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|small] [--k K] [--nb NB]

One "step" = one pass of the query hot path over one batch.
  value   : device-resident throughput (queries already in HBM), CUDA events on the launching stream
  e2e     : the same batch through the host C ABI from pinned host memory, H2D and D2H copies inside the timed region
            (sapling_b200_query_batch_bits: 2k bits per k-mer up, uint32 positions down; the whole-byte and the
            int64 entry points are timed beside it: e2e.byte_api, e2e.int64_api)
  roofline: the WHOLE STEP against the measured HBM copy bandwidth (MEASURED_PEAKS.json): the least DRAM traffic the
            step's launches can do (their streams + every index line some query touches, once) / the step time;
            `kernel` holds the dominant kernel on its own, `reference_bytes` the SURVEY 8d figure 16 + 32 (2 + P) with
            P = getLcp calls per query of the reference algorithm, counted on the device for this query set
  cpu_baseline / parity: the UNMODIFIED reference plQuery (oracle/_ref, `struct Sapling` filled member by member from the
            GPU-built parts, no constructor) on all host cores over a bounded prefix of the same queries; every answer
            is compared with the GPU's

Multi-GPU (torchrun): the index is replicated, every rank answers its own batch (weak scaling, no collective on the
query path); time = max over ranks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_G = 0x5A911C0DE5EED001
SEED_Q = 0x5A911C0DE5EED002
SEED_M = 0x5A911C0DE5EED003

WORKLOADS = {
    # name: (genome bp, queries per GPU per step, parity / cpu_baseline sample, queries per step of the reference arm,
    #        mutated queries?)
    # BASELINE.json configs[2]: the reference's constructor needs ~90 GB of host RAM and ~1 h at this size, so both CPU
    # legs fill the reference's struct from the GPU-built parts (oracle/ref_harness.cpp ref_from_parts, ~32 GB).
    "c3": (3_100_000_000, 250_000_000, 50_000_000, 25_000_000, False),
    "c2": (100_000_000, 50_000_000, 5_000_000, 5_000_000, False),
    "c1": (10_000_000, 5_000_000, 5_000_000, 5_000_000, False),
    # configs[3]: odd queries get 1-2 substitutions: the absent-k-mer path, reference behaviour F2
    "c4": (3_100_000_000, 250_000_000, 20_000_000, 10_000_000, True),
    "small": (2_000_000, 1_000_000, 500_000, 500_000, False),
}
MAXMEM = 10


def workload_name(w, k, nb):
    n, nq, _, _, mut = WORKLOADS[w]
    kind = "50% of them with 1-2 substitutions" if mut else "all present in the genome"
    return (f"{w}: synthetic {n // 1_000_000} Mbp random-ACGT genome, k={k}, maxMem={MAXMEM}, nb={'auto' if nb < 0 else nb}, "
            f"{nq // 1_000_000}M {k}-mers per GPU per step, {kind}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated, so
    that the end-to-end path copies from node-local memory (8 ranks x 2 directions of PCIe traffic otherwise cross the
    socket interconnect).  Best effort: silently keeps the inherited affinity when sysfs does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def fasta_bytes(genome: bytes, name=b"chr1", width=80):
    import numpy as np
    a = np.frombuffer(genome, dtype=np.uint8)
    full = (len(a) // width) * width
    body = np.empty((full // width, width + 1), dtype=np.uint8)
    body[:, :width] = a[:full].reshape(-1, width)
    body[:, width] = 10
    tail = a[full:].tobytes()
    return b">" + name + b"\n" + body.tobytes() + (tail + b"\n" if tail else b"")


def scratch_dir():
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize > 6 * (1 << 30):
                return tempfile.mkdtemp(prefix="sapling_bench_", dir=d)
        except Exception:
            continue
    return tempfile.mkdtemp(prefix="sapling_bench_")


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_ram_available():
    """Bytes of host RAM this process may still take: MemAvailable, capped by the cgroup limit when there is one."""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    for lim_fn, use_fn in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                           ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_fn).read().strip()
            if lim != "max" and int(lim) < (1 << 60):
                left = int(lim) - int(open(use_fn).read().strip())
                avail = left if avail is None else min(avail, left)
        except Exception:
            continue
    return avail


class stdout_to_stderr:
    """The reference's constructor prints its progress lines to stdout (C++ cout and printf): keep them off this program's
    stdout, which carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def cpu_checker_from_gpu_index(ix, k, log, threads):
    """The CPU side of parity and of both CPU timing legs: the UNMODIFIED reference `struct Sapling` with its public
    members filled from the parts the GPU built (genome, suffix array streamed in chunks, model, five error bounds);
    plQuery itself is the reference's compiled code.  Falls back to the oracle port (kind "port") when oracle/_ref is
    not built, k = 32 (reference behaviour F4: its signed k-mer arithmetic breaks) or the host has too little RAM for the
    reference's 8-byte suffix array."""
    import _oracle as O
    n = ix.n
    need_ref = 10.5 * n + 40.0 * (1 << ix.buckets) + 4e9
    avail = host_ram_available()
    t0 = time.time()
    genome = ix.reference
    xl, yl = ix.model()
    if O.ref_available() and k <= 31 and (avail is None or avail > need_ref):
        chk = O.Ref.from_parts(genome, lambda first, count: ix.rev(first, count), k, ix.buckets, xl, yl, ix.five,
                               nthreads=threads)
        kind = "reference"
    else:
        log(f"cpu checker: falling back to the oracle port (ref built: {O.ref_available()}, k={k}, host RAM available "
            f"{avail}, needed {need_ref:.3g})")
        chk = O.Port.from_parts(genome, ix.rev(), k, ix.buckets, xl, yl, ix.five)
        kind = "port"
    log(f"cpu checker ({kind}) filled from the GPU-built parts in {time.time() - t0:.1f}s")
    return chk, kind, genome


def checker_query(chk, kind, kmers, threads):
    """(answers, seconds in the query loop; string construction untimed as in sapling_example.cpp:113-140)"""
    if kind == "reference":
        return chk.query_batch(kmers, nthreads=threads, timed=True)
    return chk.query_batch_timed(kmers, nthreads=threads)


def run_reference_arm(args):
    """`--impl reference`: the reference's own plQuery timed on the box's host cores, bounded steps of the same workload.
    With a GPU in the box the index PARTS (genome, suffix array, model) come from the GPU builder -- set-up only, the
    reference's constructor would take ~1 h at 3.1 Gbp; the timed region is the unmodified header's plQuery under
    OpenMP.  Without a GPU (CPU-only containers, small workloads) the oracle's CPU builder supplies the parts."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    log = lambda m: print("[bench:reference] " + m, file=sys.stderr, flush=True)
    n, nq, _, step_q, mut = WORKLOADS[args.workload]
    k, threads = args.k, cpu_threads()
    nsteps = args.warmup + args.steps
    t0 = time.time()
    batches = []
    gpu = have_cuda()
    if gpu:
        import torch
        import sapling_b200 as S
        ix = S.Sapling.synthetic(SEED_G, n, numBuckets=args.nb, k=k, maxMem=MAXMEM, keep_host_genome=True, flags=S.QUIET)
        chk, kind, genome = cpu_checker_from_gpu_index(ix, k, log, threads)
        d = torch.empty(step_q, dtype=torch.int64, device="cuda")
        for s in range(nsteps):
            ix.sample_queries_device(SEED_Q, SEED_M if mut else 0, s * nq, step_q, d.data_ptr(), 0)
            torch.cuda.synchronize()
            batches.append(d.cpu().numpy().astype(np.uint64))
        del d
        ix.close()
    else:
        genome = O.synth_genome(SEED_G, n)
        port = O.Port.from_memory(genome, nb=args.nb, maxMem=MAXMEM, k=k)  # CPU suffix array + model (slow path, no GPU)
        if O.ref_available() and k <= 31:
            chk, kind = O.Ref.from_parts(genome, port.sa, k, port.nb, port.xlist, port.ylist, port.five,
                                         nthreads=threads), "reference"
            port.close()
        else:
            chk, kind = port, "port"
        for s in range(nsteps):
            km, _ = O.present_queries(genome, k, step_q, seed=SEED_Q + s * nq)
            if mut:
                km = O.mutate_queries(km, k, seed=SEED_M + s * nq)
            batches.append(km)
    log(f"setup: {kind} index (n={n}, k={k}, nb={chk.nb}, five={chk.five}) and {nsteps} x {step_q} queries in "
        f"{time.time() - t0:.1f}s (gpu for set-up: {gpu})")
    times = []
    for s, kmers in enumerate(batches):
        _, t = checker_query(chk, kind, kmers, threads)
        if s >= args.warmup:
            times.append(t)
    total = sum(times)
    value = step_q * len(times) / total
    line = {
        "impl": "reference", "metric": f"k={k} SA queries/sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, k, args.nb),
                   "step": f"{step_q} queries (bounded sample of the workload's {nq}-query step)",
                   "index_parts_from": "GPU builder (set-up only)" if gpu else "oracle CPU builder"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": f"{step_q} queries per step, {len(times)} steps, OpenMP over the unmodified "
                                   f"Sapling::plQuery (struct filled from parts, sapling_api.h:19-68)"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    chk.close()
    return 0


def single_thread_drivers(ix, genome, k, log):
    """the reference's own drivers as shipped (oracle/_ref/sapling_example,
    oracle/_ref/binarysearch), single thread, their own stdout timers (sapling_example.cpp:134-141,
    binarysearch.cpp:249-257).  They build through the reference's constructor, so only for genomes <= 100 Mbp."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "sapling_example")
    bs = os.path.join(ROOT, "oracle", "_ref", "binarysearch")
    if not (os.path.exists(exe) and os.path.exists(bs)):
        return None
    tmp = scratch_dir()
    out = {}
    try:
        fa = os.path.join(tmp, "g.fa")
        open(fa, "wb").write(fasta_bytes(genome))
        ix.write_sa(fa + ".sa")  # the drivers' default saFn: otherwise the reference builds the suffix array itself (sa.h DC3)
        nq = 5_000_000
        t0 = time.time()
        r = subprocess.run([exe, fa, f"k={k}", f"maxMem={MAXMEM}", f"nq={nq}", f"qLen={k}"], cwd=tmp, capture_output=True,
                           text=True, timeout=300)
        m = re.search(r"Piecewise linear time: ([0-9.eE+-]+)", r.stdout)
        c = re.search(r"Piecewise linear correctness: (\d+) out of (\d+)", r.stdout)
        if m:
            out["sapling_example"] = {"queries": nq, "seconds": float(m.group(1)), "queries_per_s": nq / float(m.group(1)),
                                      "correct": c.group(0) if c else None, "threads": 1,
                                      "wall_s_incl_build": time.time() - t0}
        r = subprocess.run([bs, fa, fa + ".sa"], cwd=tmp, capture_output=True, text=True, timeout=300)
        m = re.search(r"Binary search time: ([0-9.eE+-]+)", r.stdout)
        if m:
            out["binarysearch"] = {"queries": nq, "seconds": float(m.group(1)), "queries_per_s": nq / float(m.group(1)),
                                   "threads": 1}
    except Exception as e:
        log(f"single-thread drivers failed: {type(e).__name__}: {e}")
    finally:
        subprocess.run(["rm", "-rf", tmp])
    return out or None


def cpu_baseline_and_parity(ix, d_batch, args, sample, stream, log):
    """Rank 0, N=1: (cpu_baseline dict, parity dict) on the first `sample` queries of batch 0.  The checker is only the
    reported CPU arm and the comparison here; nothing of it is on the GPU path."""
    import numpy as np
    import torch
    import _oracle as O
    k, n, threads = args.k, ix.n, cpu_threads()
    samp = d_batch[:sample].cpu().numpy().astype(np.uint64)
    gpu_ans = torch.empty(sample, dtype=torch.int64, device="cuda")
    ix.queryBatchDevice(d_batch.data_ptr(), sample, gpu_ans.data_ptr(), stream)
    torch.cuda.synchronize()
    gpu_ans = gpu_ans.cpu().numpy()
    chk, kind, genome = cpu_checker_from_gpu_index(ix, k, log, threads)
    ref_ans, t = checker_query(chk, kind, samp, threads)
    chk.close()
    cpu = {"value": sample / t, "unit": "queries/s", "cores": threads, "kind": kind,
           "sample": f"first {sample} queries of batch 0, OpenMP over the unmodified Sapling::plQuery (struct filled from "
                     f"the GPU-built parts), string construction untimed as in sapling_example.cpp:113-140"}
    # (a predicted rank >= n is undefined in the reference -- rev[] read out of bounds, SURVEY H9 -- and not run by the harness)
    defined = ref_ans != getattr(O, "REF_UNDEFINED", None) if kind == "reference" else np.ones(sample, dtype=bool)
    parity = {"checked": int(defined.sum()), "mismatches_vs_" + kind: int(((ref_ans != gpu_ans) & defined).sum()),
              "undefined_in_reference": int((~defined).sum()),
              "minus1_answers": int((ref_ans == -1).sum()), "ranks_ge_2^31_branch": bool(n >= (1 << 31)),
              "five": list(ix.five), "nb": ix.buckets}
    if n <= 100_000_000 and kind == "reference":
        # small genomes: also the reference through its OWN constructor and files: .sap bytes, and the shipped drivers
        tmp = scratch_dir()
        fa, sa_fn, sap_fn, gpu_sap = (os.path.join(tmp, x) for x in ("g.fa", "g.fa.sa", "g.fa.sap", "gpu.sap"))
        try:
            open(fa, "wb").write(fasta_bytes(genome))
            ix.write_sa(sa_fn)
            with stdout_to_stderr():
                ref = O.Ref(fa, sa_fn, sap_fn, nb=args.nb, maxMem=MAXMEM, k=k)
            ix.write_sap(gpu_sap)
            parity["sap_bytes_identical_to_reference"] = open(gpu_sap, "rb").read() == open(sap_fn, "rb").read()
            ref.close()
        except Exception as e:
            log(f".sap byte check failed: {type(e).__name__}: {e}")
        finally:
            subprocess.run(["rm", "-rf", tmp])
        if args.workload == "c1":  # BASELINE.json configs[0]: the reference's own CPU-runnable case
            cpu["single_thread_drivers"] = single_thread_drivers(ix, genome, k, log)
    return cpu, parity


def pack_kmer_bits_device(d_kmers, kmer_bits):
    """sapling_b200.api.pack_kmer_bits on the GPU (set-up of the end-to-end input, outside every timed region): k-mer i
    goes to bits [i * kmer_bits, (i + 1) * kmer_bits) of a little-endian bit stream; eight k-mers fill kmer_bits bytes."""
    import torch
    n = d_kmers.numel()
    groups = (n + 7) // 8
    g = torch.zeros(groups * 8, dtype=torch.int64, device=d_kmers.device)
    g[:n] = d_kmers
    g = g.view(groups, 8)
    nw = (8 * kmer_bits + 63) // 64
    W = torch.zeros((groups, nw + 1), dtype=torch.int64, device=d_kmers.device)
    for j in range(8):
        w, sh = divmod(j * kmer_bits, 64)
        W[:, w] |= g[:, j] << sh
        if sh and sh + kmer_bits > 64:
            W[:, w + 1] |= (g[:, j] >> (64 - sh)) & ((1 << (sh + kmer_bits - 64)) - 1)
    stream = W.view(torch.uint8).view(groups, (nw + 1) * 8)[:, :kmer_bits].reshape(-1)
    return stream[: (n * kmer_bits + 7) // 8].contiguous()


def device_answers_host(ix, d_kmers, nq, stream):
    """Answers of the device-resident path for a batch as a host tensor (compared with the end-to-end path's)."""
    import torch
    d = torch.empty(nq, dtype=torch.int64, device="cuda")
    ix.queryBatchDevice(d_kmers.data_ptr(), nq, d.data_ptr(), stream)
    torch.cuda.synchronize()
    return d.cpu()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--nb", type=int, default=-1, help="log2 buckets; -1 = the reference's rule from maxMem")
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "none"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import sapling_b200 as S
    from sapling_b200.dist import max_over_ranks, my_shard

    rank, world, local = dist_env()
    log = (lambda m: print("[bench] " + m, file=sys.stderr, flush=True)) if rank == 0 else (lambda m: None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, nq, sample, _, mut = WORKLOADS[args.workload]
    k = args.k
    t0 = time.time()
    # k = 32 has no reference oracle (the reference's signed k-mer arithmetic breaks): the device self-check and
    # the sa_search range tests stand in
    want_cpu = (args.cpu_baseline == "auto" and rank == 0 and world == 1 and k <= 31)
    # KEEP_BUILD keeps the inverse suffix array on the device: needed to write the .sa file of the small-genome checks
    ix = S.Sapling.synthetic(SEED_G, n, numBuckets=args.nb, k=k, maxMem=MAXMEM, keep_host_genome=want_cpu,
                             flags=S.QUIET | (S.KEEP_BUILD if want_cpu and n <= 100_000_000 else 0))
    torch.cuda.synchronize()
    log(f"index built on GPU in {time.time() - t0:.1f}s: n={ix.n} k={ix.k} nb={ix.buckets} five={ix.five} "
        f"device_bytes={ix.device_bytes() / 1e6:.0f} MB")

    stream = torch.cuda.current_stream().cuda_stream
    nbatch = 3
    d_kmers = [torch.empty(nq, dtype=torch.int64, device="cuda") for _ in range(nbatch)]
    d_out = torch.empty(nq, dtype=torch.int64, device="cuda")
    for b in range(nbatch):
        # batch b of the job is a stream of world*nq queries; this rank answers its contiguous slice of it
        lo, hi = my_shard(world * nq, rank, world)
        ix.sample_queries_device(SEED_Q, SEED_M if mut else 0, b * world * nq + lo, hi - lo, d_kmers[b].data_ptr(), stream)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    for s in range(args.warmup):
        ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for s in range(args.steps):
        ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
        ev[s + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    # nvidia-smi samples every 100 ms: keep the SAME load running until about one second of it has been sampled (untimed
    # steps), so that the clocks / throttle reasons reported are a median over several samples under this load
    sustained = None
    if rank == 0 and total_ms > 0:
        extra = min(20000, max(args.steps, int(1000.0 / (total_ms / args.steps))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(extra):
            ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        sustained = {"steps": extra, "ms_per_step": e0.elapsed_time(e1) / extra,
                     "what": "the same step repeated back to back for ~1 s after the timed region (not part of `value`)"}
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(total_ms, dist, "cuda")
    value = world * nq * args.steps / (total_ms * 1e-3)
    step_mean_ms = sum(step_ms) / len(step_ms)
    kernel_name, kernel_bps = ix.query_kernel(nq)
    part_bits = ix.partition_bits(nq)
    launches0 = ix.launch_count()
    # the stages on their own: the same steps again with CUDA events recorded by the library on the launching stream
    # around each stage (a partitioned step is histogram+scans, scatter, QUERY KERNEL, un-permute)
    ix.profile(True)
    for s in range(args.steps):
        ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
    calls, stage_sum = ix.stage_ms()
    ix.profile(False)
    launches_per_step = (ix.launch_count() - launches0) // max(args.steps, 1)
    stage_ms = [x / max(calls, 1) for x in stage_sum]
    kernel_ms = stage_ms[2] if calls else step_mean_ms

    # correctness of the timed output (self-check of sapling_example.cpp:144-154, on the device)
    last = (args.steps - 1) % nbatch
    n_match, n_m1 = ix.verify_device(d_kmers[last].data_ptr(), d_out.data_ptr(), nq, stream)

    # P of SURVEY 8d: getLcp calls per query of the reference algorithm on THIS rank's queries, counted on the device
    psample = min(nq, 20_000_000)
    probes_per_q = ix.count_probes_device(d_kmers[0].data_ptr(), psample, stream) / psample

    # ---- end-to-end through the host C ABI ---------------------------------------------------
    # The host path is bound by PCIe bytes (the upload direction first: 6 of every 10 bytes), so the headline goes through
    # the densest transfer format of the C ABI (sapling_b200_query_batch_bits: a bit stream of 2k bits per k-mer up, 32-bit
    # positions down -- 9.25 bytes per query at k = 21); the whole-byte format (10 bytes) and the reference-shaped int64
    # entry point (16 bytes) are timed beside it.
    import numpy as np
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    kb = (2 * k + 7) // 8
    dev_answers = device_answers_host(ix, d_kmers[0], nq, stream)
    h_kmers = torch.empty(nq, dtype=torch.int64).pin_memory()
    h_kmers.copy_(d_kmers[0])
    torch.cuda.synchronize()
    if kb == 8:
        h_packed = h_kmers
    else:
        h_packed = torch.empty(nq * kb, dtype=torch.uint8).pin_memory()
        h_packed.copy_(torch.from_numpy(np.ascontiguousarray(h_kmers.numpy().view(np.uint8).reshape(-1, 8)[:, :kb]).reshape(-1)))
    h_bits = torch.empty((nq * 2 * k + 7) // 8, dtype=torch.uint8).pin_memory()
    h_bits.copy_(pack_kmer_bits_device(d_kmers[0], 2 * k))
    torch.cuda.synchronize()
    h_out32 = torch.empty(nq, dtype=torch.int32).pin_memory()
    h_out = torch.empty(nq, dtype=torch.int64).pin_memory()

    def timed(fn):
        fn()  # warm-up (allocates staging)
        fn()
        barrier()
        l0 = ix.launch_count()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, dist, "cuda")
        return world * nq * e2e_steps / dt, ix.launch_count() - l0

    def answers32():
        got32 = h_out32.to(torch.int64) & 0xFFFFFFFF
        return torch.where(got32 == 0xFFFFFFFF, torch.full_like(got32, -1), got32)

    e2e_bytes_value, _ = timed(lambda: ix.queryBatchU32(h_packed, kmer_bytes=kb, out=h_out32, nq=nq))
    e2e_equal = bool(torch.equal(answers32(), dev_answers))
    h_out32.zero_()
    e2e_value, e2e_launches = timed(lambda: ix.queryBatchBits(h_bits, 2 * k, nq, out=h_out32))
    e2e64_value, _ = timed(lambda: ix.queryBatch(h_kmers, out=h_out))
    e2e_equal = bool(e2e_equal and torch.equal(answers32(), dev_answers) and torch.equal(h_out, dev_answers))
    h2d_bits = int(h_bits.numel())
    del h_kmers, h_out, h_packed, h_bits, h_out32, dev_answers

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline + CPU baseline (rank 0) ----------------------------------------------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    peak, peak_src = peaks()
    cpu, parity = None, None
    if want_cpu:
        try:
            cpu, parity = cpu_baseline_and_parity(ix, d_kmers[0], args, sample, stream, log)
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            log(f"cpu_baseline failed: {type(e).__name__}: {e}")

    ref_bytes_per_query = 16 + 32 * (2 + probes_per_q)
    # Least DRAM traffic of the dominant kernel and of the whole step.  Partitioned step: the streams of
    # passes A, B, Q, U (8 + 16 + 16 + 16 bytes per query) plus every index line that at least one query of the batch
    # touches, once (rank lines: 128 B per 16 ranks; narrow model: 8 B per bucket; Poisson coverage).  Unpartitioned:
    # nothing is shared, a query pays the reference's sectors.
    if part_bits:
        lines = n / 16.0
        line_bytes = 128.0 * lines * (1.0 - math.exp(-nq / lines))
        buckets = float(1 << ix.buckets)
        model_bytes = 8.0 * buckets * (1.0 - math.exp(-nq / buckets))
        kernel_bytes = 16.0 * nq + line_bytes + model_bytes
        step_bytes = 40.0 * nq + kernel_bytes
    else:
        kernel_bytes = step_bytes = ref_bytes_per_query * nq
    step_gbs = step_bytes / (step_mean_ms * 1e-3) / 1e9
    kernel_gbs = kernel_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel on this workload (dram__bytes_read.sum + dram__bytes_write.sum of the
    # committed ncu --set full capture of the same kernel; not measurable without the profiler)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and k == 21 and args.nb == -1:
        try:
            t = json.load(open(tp)).get(f"{args.workload}:{kernel_name}")
            if t:
                traffic, traffic_src = float(t["per_launch_bytes"]), t.get("source")
        except Exception:
            traffic = None
    gather = None
    try:
        gather = S.gather_bench(12 << 30, 1 << 28, 3)
    except Exception as e:
        log(f"gather bench failed: {e}")

    line = {
        "metric": f"k={k} SA queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, k, args.nb), "queries_per_step_per_gpu": nq,
                   "index": "replicated per GPU", "nb": ix.buckets, "error_bounds": list(ix.five),
                   "l2": f"inputs larger than L2 ({nq * 8 // 1_000_000} MB k-mers + {nq * 8 // 1_000_000} MB results streamed "
                         f"per step, 3 rotating batches; {ix.device_bytes() // 1_000_000} MB index gathered)",
                   "seeds": {"genome": hex(SEED_G), "queries": hex(SEED_Q), "mutations": hex(SEED_M) if mut else None}},
        "roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "what": "whole step (all launches): least DRAM bytes the step can move / step time",
                     "bytes_per_query": step_bytes / nq, "step_ms": step_mean_ms, "partition_bits": part_bits,
                     "stage_ms": dict(zip(("hist_scan", "scatter", "query_kernel", "unpermute"), stage_ms)),
                     "kernel": {"name": kernel_name, "blocks_per_sm": kernel_bps, "ms": kernel_ms,
                                "share_of_step": kernel_ms / step_mean_ms if step_mean_ms else None,
                                "bytes_per_query": kernel_bytes / nq, "achieved": kernel_gbs, "frac": kernel_gbs / peak,
                                "dram_gbs_from_traffic": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None},
                     "reference_bytes": {"bytes_per_query": ref_bytes_per_query, "probes_per_query": probes_per_q,
                                         "probes_source": f"getLcp calls of the reference algorithm counted on the device "
                                                          f"over this rank's first {psample} queries",
                                         "achieved": nq * ref_bytes_per_query / (kernel_ms * 1e-3) / 1e9,
                                         "frac": nq * ref_bytes_per_query / (kernel_ms * 1e-3) / 1e9 / peak,
                                         "what": "SURVEY 8d: what the REFERENCE's access pattern would move per query / "
                                                 "the dominant kernel's time; above 1 only says the kernel does not move "
                                                 "those bytes (a partitioned batch shares index lines)"},
                     "random_sector_gather_gbs": gather},
        "sustained": sustained,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d_bits, "d2h_bytes_per_step": nq * 4,
                "steps": e2e_steps, "api": f"sapling_b200_query_batch_bits (pinned host buffers: a bit stream of {2 * k} bits "
                                           f"per k-mer up, uint32 positions down)", "numa_node": numa_node,
                "byte_api": {"value": e2e_bytes_value, "api": f"sapling_b200_query_batch_u32 ({kb}-byte k-mers up, uint32 "
                                                              f"positions down)",
                             "h2d_bytes_per_step": nq * kb, "d2h_bytes_per_step": nq * 4},
                "int64_api": {"value": e2e64_value, "api": "sapling_b200_query_batch (uint64 k-mers up, int64 answers down)",
                              "h2d_bytes_per_step": nq * 8, "d2h_bytes_per_step": nq * 8},
                "answers_equal_device_path": e2e_equal},
        "gpu_launches": args.steps * launches_per_step, "e2e_gpu_launches": e2e_launches,
        "clocks": clocks, "self_check": {"matching": int(n_match), "minus1": int(n_m1), "of": nq},
        "parity": parity,
    }
    print(json.dumps(line), flush=True)
    ix.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

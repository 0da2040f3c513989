#!/usr/bin/env python
"""bench.py -- k=21 SAPLING suffix-array queries/sec on B200.

Workload (BASELINE.json configs[1], "c2"): synthetic 100 Mbp random-ACGT genome (counter-based
generator), suffix array and .sap model built on the GPU with the reference's default
parameters (k=21, maxMem=10 -> nb=23), 50 M 21-mers sampled from the genome per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3|small]

One "step" = one pass of the query hot path over one batch (50 M queries at c2).
  value  : device-resident throughput (queries already in HBM), CUDA events on the launching stream
  e2e    : the same batch through the host C ABI (sapling_b200_query_batch) from pinned host memory,
           H2D and D2H copies inside the timed region
  roofline: algorithmic bytes/query = 16 + 32*(2 + P) (P = genome probes/query executed
           by the reference algorithm on this query set, counted by the oracle) / kernel time,
           against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline: the reference's own plQuery (oracle/_ref, unmodified sapling_api.h) on all host
           cores over a bounded prefix of the same queries; its answers are also compared with the
           GPU's (bit-exact parity at full size)

Multi-GPU (torchrun): the index is replicated, every rank answers its own batch (weak scaling, no
collective on the query path); time = max over ranks.
"""
import argparse
import json
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

WORKLOADS = {
    # name: (genome bp, queries per step, cpu sample)
    "c2": (100_000_000, 50_000_000, 5_000_000),
    "c1": (10_000_000, 5_000_000, 5_000_000),
    # BASELINE.json configs[2]: human-scale genome; 250 M queries per GPU per step (4 GPU-steps make the 1 B of the
    # config).  The reference itself needs ~90 GB of host RAM and ~1 h to construct at this size, so the CPU baseline
    # and the parity check use the oracle port built from the GPU index's parts.
    "c3": (3_100_000_000, 250_000_000, 2_000_000),
    "small": (2_000_000, 1_000_000, 500_000),
}
K = 21
MAXMEM = 10


def workload_name(w):
    n, nq, _ = WORKLOADS[w]
    return (f"{w}: synthetic {n // 1_000_000} Mbp random-ACGT genome, k={K}, maxMem={MAXMEM}, "
            f"{nq // 1_000_000}M present 21-mers per GPU per step")


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def build_reference_index(workload, log):
    """The unmodified reference (oracle/_ref) constructed through its own constructor from a FASTA and a .sa file.
    The suffix array (unique for a text) is produced by the GPU builder when a GPU is present, else by the oracle's
    CPU builder; the .sap model is built by the reference itself."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import _oracle as O
    n, nq, sample = WORKLOADS[workload]
    tmp = scratch_dir()
    fa, sa_fn, sap_fn = (os.path.join(tmp, "g.fa"), os.path.join(tmp, "g.fa.sa"), os.path.join(tmp, "g.fa.sap"))
    t0 = time.time()
    gpu_ix = None
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        import sapling_b200 as S
        gpu_ix = S.Sapling.synthetic(SEED_G, n, k=K, maxMem=MAXMEM, keep_host_genome=True, flags=S.QUIET | S.KEEP_BUILD)
        genome = gpu_ix.reference
        gpu_ix.write_sa(sa_fn)
    else:
        genome = O.synth_genome(SEED_G, n)
        port = O.Port.from_memory(genome, k=K)  # CPU suffix array (slow path, no GPU)
        port.write_sa(sa_fn)
        port.close()
    open(fa, "wb").write(fasta_bytes(genome))
    log(f"setup: genome+.sa written to {tmp} in {time.time() - t0:.1f}s (gpu={have_gpu})")
    t0 = time.time()
    kind = "reference" if O.ref_available() else "port"
    if kind == "reference":
        ref = O.Ref(fa, sa_fn, sap_fn, nb=-1, maxMem=MAXMEM, k=K)
    else:
        ref = O.Port.open(fa, sa_fn, sap_fn, nb=-1, maxMem=MAXMEM, k=K)
    log(f"setup: {kind} index constructed (its own .sap build) in {time.time() - t0:.1f}s; nb={ref.nb} five={ref.five}")
    return ref, kind, genome, gpu_ix, tmp, (fa, sa_fn, sap_fn)


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    log = lambda m: print("[bench:reference] " + m, file=sys.stderr, flush=True)
    n, nq, sample = WORKLOADS[args.workload]
    ref, kind, genome, gpu_ix, tmp, files = build_reference_index(args.workload, log)
    if gpu_ix is not None:
        gpu_ix.close()
    threads = cpu_threads()
    batches = []
    for s in range(args.warmup + args.steps):
        kmers, _ = O.present_queries(genome, K, sample, seed=SEED_Q + s * nq)
        batches.append(kmers)
    times = []
    for s, kmers in enumerate(batches):
        if kind == "reference":
            _, t = ref.query_batch(kmers, nthreads=threads, timed=True)
        else:
            _, t = ref.query_batch_timed(kmers, nthreads=threads)
        if s >= args.warmup:
            times.append(t)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": "k=21 SA queries/sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "step": f"{sample} queries (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} present 21-mers per step, {len(times)} steps, OpenMP over Sapling::plQuery"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    ref.close()
    for f in files:
        try:
            os.remove(f)
        except OSError:
            pass
    try:
        os.rmdir(tmp)
    except OSError:
        pass
    return 0


def cpu_baseline_and_parity(ix, d_batch, n, sample, stream, log):
    """Rank 0, N=1: (probes/query counted by the oracle, cpu_baseline dict, parity dict) on the first `sample` queries of
    batch 0.  The oracle is only the checker / the reported CPU arm here; nothing of it is on the GPU path."""
    import numpy as np
    import torch
    import _oracle as O
    genome = ix.reference
    samp = d_batch[:sample].cpu().numpy().astype(np.uint64)
    gpu_ans = torch.empty(sample, dtype=torch.int64, device="cuda")
    ix.queryBatchDevice(d_batch.data_ptr(), sample, gpu_ans.data_ptr(), stream)
    torch.cuda.synchronize()
    gpu_ans = gpu_ans.cpu().numpy()
    threads = cpu_threads()
    # P: probes/query of the reference algorithm, counted by the oracle port on this sample
    xl, yl = ix.model()
    port = O.Port.from_parts(genome, ix.rev(), K, ix.buckets, xl, yl, ix.five)
    psamp = samp[:min(sample, 2_000_000)]
    pans, ptot, _ = port.query_batch(psamp, nthreads=threads, stats=True)
    probes_per_q = ptot / len(psamp)
    port_equal = bool(np.array_equal(pans, gpu_ans[:len(psamp)]))
    if n >= (1 << 31):
        # c3: the oracle port is the CPU arm (the reference needs ~90 GB and ~1 h to construct at this size)
        _, t = port.query_batch_timed(samp, nthreads=threads)
        port.close()
        cpu = {"value": sample / t, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} queries of batch 0 (present 21-mers), OpenMP over the oracle's plQuery "
                         f"restatement on the GPU-built index parts, string construction untimed"}
        parity = {"checked": int(len(psamp)), "mismatches_vs_port": int((pans != gpu_ans[:len(psamp)]).sum()),
                  "oracle_port_equal": port_equal, "five": list(ix.five), "nb": ix.buckets}
        return probes_per_q, cpu, parity
    port.close()
    # the reference itself, through its own constructor and files
    tmp = scratch_dir()
    fa, sa_fn, sap_fn = (os.path.join(tmp, "g.fa"), os.path.join(tmp, "g.fa.sa"), os.path.join(tmp, "g.fa.sap"))
    gpu_sap = os.path.join(tmp, "gpu.sap")
    try:
        t0 = time.time()
        open(fa, "wb").write(fasta_bytes(genome))
        ix.write_sa(sa_fn)
        kind = "reference" if O.ref_available() else "port"
        if kind == "reference":
            ref = O.Ref(fa, sa_fn, sap_fn, nb=-1, maxMem=MAXMEM, k=K)
        else:
            ref = O.Port.open(fa, sa_fn, sap_fn, nb=-1, maxMem=MAXMEM, k=K)
        log(f"cpu_baseline: {kind} constructed from files in {time.time() - t0:.1f}s")
        ix.write_sap(gpu_sap)
        sap_identical = open(gpu_sap, "rb").read() == open(sap_fn, "rb").read()
        if kind == "reference":
            ref_ans, t = ref.query_batch(samp, nthreads=threads, timed=True)
        else:
            ref_ans, t = ref.query_batch_timed(samp, nthreads=threads)
        ref.close()
    finally:
        for f in (fa, sa_fn, sap_fn, gpu_sap):
            try:
                os.remove(f)
            except OSError:
                pass
        try:
            os.rmdir(tmp)
        except OSError:
            pass
    cpu = {"value": sample / t, "unit": "queries/s", "cores": threads, "kind": kind,
           "sample": f"first {sample} queries of batch 0 (present 21-mers), OpenMP over Sapling::plQuery, "
                     f"string construction untimed as in sapling_example.cpp:113-140"}
    parity = {"checked": int(sample), "mismatches_vs_" + kind: int((ref_ans != gpu_ans).sum()),
              "oracle_port_equal": port_equal, "sap_bytes_identical_to_" + kind: bool(sap_identical),
              "five": list(ix.five), "nb": ix.buckets}
    return probes_per_q, cpu, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "none"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
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

    n, nq, sample = WORKLOADS[args.workload]
    t0 = time.time()
    want_cpu = (args.cpu_baseline == "auto" and rank == 0 and world == 1)
    # KEEP_BUILD keeps the inverse suffix array on the device: needed to write the .sa file the reference arm reads
    ix = S.Sapling.synthetic(SEED_G, n, k=K, maxMem=MAXMEM, keep_host_genome=want_cpu,
                             flags=S.QUIET | (S.KEEP_BUILD if want_cpu and n < (1 << 31) else 0))
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
        ix.sample_queries_device(SEED_Q, 0, b * world * nq + lo, hi - lo, d_kmers[b].data_ptr(), stream)
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
    # nvidia-smi samples every 100 ms and the timed region is tens of milliseconds long: keep the SAME load running for about
    # half a second more (untimed) so that the clocks / throttle reasons reported are a median over several samples
    if rank == 0 and total_ms > 0:
        extra = min(20000, max(args.steps, int(500.0 / (total_ms / args.steps))))
        for s in range(extra):
            ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(total_ms, dist, "cuda")
    value = world * nq * args.steps / (total_ms * 1e-3)
    step_mean_ms = sum(step_ms) / len(step_ms)
    kernel_name, kernel_bps = ix.query_kernel(nq)
    part_bits = ix.partition_bits(nq)
    launches_per_step = 8 if part_bits else 1  # hist, 3 column-scan passes, bin scan, scatter, QUERY, un-permute
    # the query kernel on its own: the same steps again with CUDA events recorded by the library on the launching
    # stream around each stage (a partitioned step is histogram+scans, scatter, QUERY KERNEL, un-permute)
    ix.profile(True)
    for s in range(args.steps):
        ix.queryBatchDevice(d_kmers[s % nbatch].data_ptr(), nq, d_out.data_ptr(), stream)
    calls, stage_sum = ix.stage_ms()
    ix.profile(False)
    stage_ms = [x / max(calls, 1) for x in stage_sum]
    kernel_ms = stage_ms[2] if calls else step_mean_ms

    # correctness of the timed output (self-check of sapling_example.cpp:144-154, on the device)
    last = (args.steps - 1) % nbatch
    n_match, n_m1 = ix.verify_device(d_kmers[last].data_ptr(), d_out.data_ptr(), nq, stream)

    # ---- end-to-end through the host C ABI ---------------------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    h_kmers = torch.empty(nq, dtype=torch.int64).pin_memory()
    h_out = torch.empty(nq, dtype=torch.int64).pin_memory()
    h_kmers.copy_(d_kmers[0])
    torch.cuda.synchronize()
    ix.queryBatch(h_kmers, out=h_out)  # warm-up (allocates staging)
    ix.queryBatch(h_kmers, out=h_out)
    barrier()
    launches0 = ix.launch_count()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ix.queryBatch(h_kmers, out=h_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s, dist, "cuda")
    e2e_value = world * nq * e2e_steps / e2e_s
    e2e_launches = ix.launch_count() - launches0

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline + CPU baseline (rank 0) ----------------------------------------------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    peak, peak_src = peaks()
    probes_per_q, cpu, parity = None, None, None
    if want_cpu:
        try:
            probes_per_q, cpu, parity = cpu_baseline_and_parity(ix, d_kmers[0], n, sample, stream, log)
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            log(f"cpu_baseline failed: {type(e).__name__}: {e}")

    if probes_per_q is None:
        probes_per_q = {"c2": 2.34, "c1": 2.47}.get(args.workload, 2.4)  # SURVEY 3.3 [probe]
        p_src = "survey value"
    else:
        p_src = "counted by the oracle on this query set"
    bytes_per_query = 16 + 32 * (2 + probes_per_q)
    achieved = nq * bytes_per_query / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes per launch of THIS kernel on THIS workload (dram__bytes_read.sum + dram__bytes_write.sum of the committed
    # ncu --set full capture)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp)).get(f"{args.workload}:{kernel_name}")
            if t:
                traffic, traffic_src = float(t["per_launch_bytes"]), t.get("source")
        except Exception:
            traffic = None
    # Whole-step HBM bound of the partitioned pipeline: what one step has to move at the least -- the streams
    # of passes A, B, Q, U (8 + 16 + 16 + 16 bytes per query) plus every index line that at least one query of the batch
    # touches, once (rank lines 128 B per 16 ranks and the narrow model 8 B per bucket, Poisson coverage).  Unpartitioned: the
    # per-query figure above, there is no sharing.
    import math
    if part_bits:
        lines = n / 16.0
        line_bytes = 128.0 * lines * (1.0 - math.exp(-nq / lines))
        buckets = float(1 << ix.buckets)
        model_bytes = 8.0 * buckets * (1.0 - math.exp(-nq / buckets))
        step_bytes = 56.0 * nq + line_bytes + model_bytes
    else:
        step_bytes = bytes_per_query * nq
    step_gbs = step_bytes / (step_mean_ms * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "bytes_per_query": step_bytes / nq, "step_ms": step_mean_ms,
                     "what": "minimum DRAM bytes of one whole step (all launches) / step time"}
    gather = None
    try:
        gather = S.gather_bench(12 << 30, 1 << 28, 3)
    except Exception as e:
        log(f"gather bench failed: {e}")

    line = {
        "metric": "k=21 SA queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "queries_per_step_per_gpu": nq, "index": "replicated per GPU",
                   "nb": ix.buckets, "error_bounds": list(ix.five), "l2": f"inputs larger than L2 ({nq * 8 // 1_000_000} MB k-mers + "
                   f"{nq * 8 // 1_000_000} MB results streamed per step, 3 rotating batches; "
                   f"{ix.device_bytes() // 1_000_000} MB index gathered)",
                   "seeds": {"genome": hex(SEED_G), "queries": hex(SEED_Q)}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "dram_gbs_from_traffic": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                     "note": ("achieved counts the REFERENCE algorithm's bytes per query (SURVEY 8d); a partitioned batch "
                              "shares index lines between queries, so the kernel's own DRAM traffic is lower and the kernel "
                              "is bound by instruction issue (DESIGN.md 4.2)") if part_bits else None,
                     "peak_source": peak_src, "bytes_per_query": bytes_per_query,
                     "probes_per_query": probes_per_q, "probes_source": p_src, "kernel": kernel_name,
                     "blocks_per_sm": kernel_bps, "kernel_ms": kernel_ms, "step_ms": step_mean_ms,
                     "kernel_share_of_step": kernel_ms / step_mean_ms if step_mean_ms else None,
                     "partition_bits": part_bits,
                     "stage_ms": dict(zip(("hist_scan", "scatter", "query_kernel", "unpermute"), stage_ms)),
                     "random_sector_gather_gbs": gather},
        "roofline_step": roofline_step,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * 8, "d2h_bytes_per_step": nq * 8,
                "steps": e2e_steps, "api": "sapling_b200_query_batch (pinned host buffers)", "numa_node": numa_node},
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

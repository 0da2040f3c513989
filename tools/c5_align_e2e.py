#!/usr/bin/env python
"""BASELINE.json configs[4] end to end: the whole aligner.  Synthetic reads (150 bp, 1 % substitutions, a few N, a third
reverse-complemented) written as FASTQ, then
  align_ref   = the reference's align.cpp (with its one-line fix, oracle/ref_align_harness.cpp), single thread as shipped
  align_b200  = sapling_b200/host/align_b200.cpp: seed lookups batched on the GPU, SSW extension on all host cores
on the same genome and index files; wall clock of each (index load included and reported apart via the drivers' own
lines), and whether the SAM bodies are identical.

  python tools/c5_align_e2e.py [n=1e8] [n_reads=1e6] [ref_reads=1e5]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import sapling_b200 as S  # noqa: E402
import _oracle as O  # noqa: E402
from c5_align_seeds import make_reads, SEED_G, K  # noqa: E402

ALIGN_REF = os.path.join(ROOT, "oracle", "_ref", "align_ref")
ALIGN_B200 = os.path.join(ROOT, "sapling_b200", "bin", "align_b200")


def write_fastq(path, reads):
    n, L = reads.shape
    rec = np.empty((n, 0), dtype=np.uint8)
    names = [b"@r%d\n" % i for i in range(n)]
    with open(path, "wb") as f:
        qual = b"I" * L
        for i in range(n):
            f.write(names[i] + reads[i].tobytes() + b"\n+\n" + qual + b"\n")


def run(cmd, cwd):
    t0 = time.time()
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    return time.time() - t0, p


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
    ref_reads = int(float(sys.argv[3])) if len(sys.argv) > 3 else 100_000
    tmp = tempfile.mkdtemp(prefix="c5_", dir="/dev/shm")
    res = {"n": n, "k": K, "n_reads": n_reads, "ref_reads": ref_reads, "cores": os.cpu_count()}
    try:
        ix = S.Sapling.synthetic(SEED_G, n, k=K, maxMem=10, keep_host_genome=True, flags=S.QUIET | S.KEEP_BUILD)
        genome = ix.reference
        fa = os.path.join(tmp, "g.fa")
        O.write_fasta(fa, genome)
        ix.write_sa(fa + ".sa")
        ix.write_sap(fa + f"_k{K}.sap")   # align.cpp:168 names the model file <ref>_k<k>.sap
        ix.close()
        reads, _ = make_reads(genome, n_reads)
        fq_all, fq_ref = os.path.join(tmp, "all.fq"), os.path.join(tmp, "ref.fq")
        write_fastq(fq_all, reads)
        write_fastq(fq_ref, reads[:ref_reads])
        # the GPU-seeded aligner: all reads, and the reference's subset
        t, p = run([ALIGN_B200, fq_all, fa, os.path.join(tmp, "b200_all.sam")], tmp)
        res["align_b200"] = {"reads": n_reads, "wall_s": round(t, 2), "reads_per_s": round(n_reads / t, 1),
                             "report": p.stdout.strip().splitlines()[-1] if p.stdout else p.stderr[-300:]}
        t, p = run([ALIGN_B200, fq_ref, fa, os.path.join(tmp, "b200_ref.sam")], tmp)
        res["align_b200_subset"] = {"reads": ref_reads, "wall_s": round(t, 2)}
        if os.path.exists(ALIGN_REF):
            t, p = run([ALIGN_REF, fq_ref, fa, os.path.join(tmp, "ref.sam")], tmp)
            res["align_ref"] = {"reads": ref_reads, "wall_s": round(t, 2), "reads_per_s": round(ref_reads / t, 1), "threads": 1,
                                "rc": p.returncode}
            body = lambda fn: [l for l in open(fn) if not l.startswith("@PG")]
            res["sam_identical_on_subset"] = body(os.path.join(tmp, "ref.sam")) == body(os.path.join(tmp, "b200_ref.sam"))
            res["speedup_reads_per_s"] = round(res["align_b200"]["reads_per_s"] / res["align_ref"]["reads_per_s"], 1)
    finally:
        subprocess.run(["rm", "-rf", tmp])
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "c5_align_e2e.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

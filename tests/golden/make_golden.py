#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsapling_ref.so, i.e.
/root/reference/src/sapling_api.h compiled as is).  Runs only where /root/reference exists.

Each fixture pins, for one genome x (k, nb): the cleaned genome, the suffix array, the model the
reference's buildPiecewiseLinear produced (xlist, ylist, five error bounds, perfectPredictions),
and the reference's plQuery answers for present / mutated / random k-mers and for variable-length
strings (the sapling_example sweep k-10..k+80).  Queries on which the reference itself has
undefined behaviour (predicted rank >= n -> out-of-bounds rev[]) are excluded and counted.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _fixtures as F  # noqa: E402
import _oracle as O  # noqa: E402

DEMO = "/root/reference/Complete-Striped-Smith-Waterman-Library/demo"

CASES = [
    # (fixture genome, k, nb)
    ("rand20k", 21, -1), ("rand20k", 11, 4), ("rand20k", 31, 12), ("rand20k", 16, 8),
    ("gc1991", 21, -1), ("gc0110", 16, 10), ("polyC", 21, -1), ("tandem_CT", 21, 8),
    ("tandem_ACGTCGTAGTACTACG", 11, 4), ("tandem50", 21, 8), ("repeat_tailA", 21, -1),
    ("ssw100k_slice", 21, -1), ("ssw100k_slice", 16, 6),
]


def genomes():
    g = F.small_genomes()
    p = os.path.join(DEMO, "100k.fa")
    real, _ = O.clean_fasta_text(open(p, "rb").read())
    g["ssw100k_slice"] = real[20000:50000]  # real human sequence shipped with the reference (SSW demo data)
    return g


def main():
    O.build()
    G = genomes()
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        for name, k, nb in CASES:
            g = G[name]
            n = len(g)
            fa = os.path.join(tmp, f"{name}.fa")
            O.write_fasta(fa, g)
            tag = f"{name}.k{k}.nb{nb}"
            ref = O.Ref(fa, os.path.join(tmp, name + ".sa"), os.path.join(tmp, tag + ".sap"), nb=nb, k=k)
            port = O.Port.from_memory(g, nb=nb, k=k)  # only used to find the reference-UB queries
            kmers = F.query_mix(g, k, 1500)
            ok = np.array([port.predict(int(x)) < n for x in kmers], dtype=bool)
            kmers = kmers[ok]
            answers = ref.query_batch(kmers)
            strs = [s for s in F.var_len_strings(g, k, 12)]
            km = [ref.kmerize_adjusted(len(s), s) for s in strs]
            keep = []
            for s, x in zip(strs, km):
                _, _, fl = port.query_str(s, x, want_probes=True)
                keep.append(fl == 0)
            strs = [s for s, kp in zip(strs, keep) if kp]
            km = [x for x, kp in zip(km, keep) if kp]
            sans = np.array([ref.query_str(s, x) for s, x in zip(strs, km)], dtype=np.int64)
            out = os.path.join(HERE, tag + ".npz")
            np.savez_compressed(
                out, genome=np.frombuffer(g, dtype=np.uint8), k=k, nb_arg=nb, nb=ref.nb, five=np.array(ref.five),
                perfect=ref.perfect, sa=ref.sa.astype(np.uint32), xlist=ref.xlist, ylist=ref.ylist,
                kmers=kmers, answers=answers, excluded_ub=int((~ok).sum()),
                strings=np.frombuffer(b"".join(strs), dtype=np.uint8),
                string_lens=np.array([len(s) for s in strs], dtype=np.uint32),
                string_kmers=np.array(km, dtype=np.int64), string_answers=sans)
            print(f"{tag}: n={n} nb={ref.nb} five={ref.five} kmers={len(kmers)} strings={len(strs)} "
                  f"-> {os.path.getsize(out) // 1024} KiB")
            ref.close()
            port.close()


SEED_CASES = [("rand20k", 16), ("gc1991", 16), ("tandem50", 16), ("ssw100k_slice", 16), ("repeat_tailA", 21)]


def main_seeds():
    """seeds.<genome>.k<k>.npz: the align.cpp:259-300 seed loop driven through the unmodified reference's own methods
    (oracle/ref_harness.cpp ref_seed_batch) on simulated reads; `defined` masks the slots on which the reference itself
    reads out of bounds (countHitsLeft on the last rank; a seed whose predicted rank is >= n, which the harness does not run:
    ref_pos == -2)."""
    G = genomes()
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        for name, k in SEED_CASES:
            g = G[name]
            fa = os.path.join(tmp, f"{name}.fa")
            O.write_fasta(fa, g)
            ref = O.Ref(fa, os.path.join(tmp, name + ".sa"), os.path.join(tmp, f"{name}.k{k}.sap"), k=k)
            reads, _ = O.simulate_reads(g, 150, min(150, len(g) // 4))
            reads += [b"ACGT", g[100:100 + k], g[7:7 + k + 1], b"N" * 60, g[5:155].lower(), g[-150:], g[:150]]
            rp, sp, lf, rt = ref.seed_batch(reads, 7, 32)
            defined = ~((rp >= 0) & (sp == len(g) - 1)) & (rp != -2)
            out = os.path.join(HERE, f"seeds.{name}.k{k}.npz")
            np.savez_compressed(out, genome=np.frombuffer(g, dtype=np.uint8), k=k, num_seeds=7, max_hits=32,
                                reads=np.frombuffer(b"".join(reads), dtype=np.uint8),
                                read_lens=np.array([len(r) for r in reads], dtype=np.uint32),
                                ref_pos=rp, sa_pos=sp, left=lf, right=rt, defined=defined)
            print(f"seeds.{name}.k{k}: reads={len(reads)} hits={(rp >= 0).sum()} of {rp.size} "
                  f"max left/right={lf.max()}/{rt.max()} undefined={(~defined).sum()} -> {os.path.getsize(out) // 1024} KiB")
            ref.close()


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "seeds":
        main()
    main_seeds()

#!/usr/bin/env python
"""Pin oracle/sapling_oracle.c against the unmodified reference (oracle/_ref/libsapling_ref.so).

Runs only where /root/reference exists (this container).  For every fixture genome x (k, nb):
builds the index with BOTH implementations (reference: its own DC3 + buildPiecewiseLinear), then
compares .sa bytes, .sap bytes, predictions and plQuery results on present, mutated and
variable-length queries.  Prints one line per case and exits non-zero on any difference.

    python tools/pin_oracle.py [--big]
"""
import argparse
import filecmp
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _oracle as O  # noqa: E402

DEMO = "/root/reference/Complete-Striped-Smith-Waterman-Library/demo"


def fixture_genomes(big):
    rng = np.random.default_rng(12345)

    def freq(n, w):
        w = np.array(w, dtype=float)
        return bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=n, p=w / w.sum())])

    out = []
    out.append(("rand20k", O.synth_genome(O.SEED_G, 20000)))
    out.append(("rand300k", O.synth_genome(O.SEED_G + 7, 300000)))
    # recipes restated from eval/SuffixArraySim/SuffixArraySimulatedSequences.java:13-26
    out.append(("gc2332", freq(50000, [2, 3, 3, 2])))
    out.append(("gc1991", freq(50000, [1, 9, 9, 1])))
    out.append(("gc0110", freq(30000, [0, 1, 1, 0])))
    out.append(("polyC", b"C" * 5000))
    for unit in ("CT", "CAT", "ACGT", "ACTTCA", "ACGTCGTAGTACTACG"):
        out.append(("tandem_" + unit, (unit * (20000 // len(unit) + 1)).encode()[:20000]))
    fifty = freq(50, [1, 1, 1, 1])
    out.append(("tandem50", (fifty * 400)[:20000]))
    # real sequence from the SSW demo data shipped in the reference tree
    for nm in ("100k.fa", "1M.fa") + (("10M.fa",) if big else ()):
        p = os.path.join(DEMO, nm)
        if os.path.exists(p):
            g, _ = O.clean_fasta_text(open(p, "rb").read())
            out.append(("ssw_" + nm, g))
    if big:
        out.append(("rand5M", O.synth_genome(O.SEED_G + 11, 5_000_000)))
    return out


def var_len_queries(genome, k, count, rng):
    """(string, kmerizeAdjusted, length) for the sapling_example sweep k-10..k+80 (sapling_example.cpp:93-98)"""
    qs = []
    n = len(genome)
    for L in (max(1, k - 10), k, k + 10, k + 20, k + 30, k + 80):
        if L >= n:
            continue
        for p in rng.integers(0, n - L, size=count):
            s = genome[p:p + L]
            qs.append((s, L))
    return qs


def run_case(name, genome, k, nb, tmp, nq=20000):
    fa = os.path.join(tmp, f"{name}.fa")
    O.write_fasta(fa, genome)
    tag = f"{name}.k{k}.nb{nb}"
    sa_r, sap_r = os.path.join(tmp, name + ".ref.sa"), os.path.join(tmp, tag + ".ref.sap")
    sa_p, sap_p = os.path.join(tmp, name + ".port.sa"), os.path.join(tmp, tag + ".port.sap")
    ref = O.Ref(fa, sa_r, sap_r, nb=nb, maxMem=-1, k=k)
    port = O.Port.open(fa, sa_p, sap_p, nb=nb, maxMem=-1, k=k)
    bad = []
    if not filecmp.cmp(sa_r, sa_p, shallow=False):
        bad.append(".sa bytes")
    if not filecmp.cmp(sap_r, sap_p, shallow=False):
        bad.append(".sap bytes")
    if ref.five != port.five or ref.nb != port.nb or ref.perfect != port.perfect:
        bad.append(f"stats ref={ref.five},{ref.nb},{ref.perfect} port={port.five},{port.nb},{port.perfect}")
    n = len(genome)
    nq = min(nq, 4 * n)
    present, _ = O.present_queries(genome, k, nq)
    mutated = O.mutate_queries(present, k)
    rng = np.random.default_rng(99)
    randq = rng.integers(0, 1 << (2 * k), size=nq // 4, dtype=np.uint64)
    kmers = np.concatenate([present, mutated, randq])
    # predicted >= n is undefined in the reference (out-of-bounds rev[]): exclude and count
    pr, probes, oob = port.query_batch(kmers, nthreads=4, stats=True)
    pred_ok = np.array([port.predict(int(x)) < n for x in kmers[len(present):]], dtype=bool)
    keep = np.concatenate([np.ones(len(present), dtype=bool), pred_ok])
    rr = ref.query_batch(kmers[keep], nthreads=4)
    if not np.array_equal(rr, pr[keep]):
        d = np.flatnonzero(rr != pr[keep])
        bad.append(f"plQuery batch: {len(d)} diffs, first kmer={int(kmers[keep][d[0]])} ref={rr[d[0]]} port={pr[keep][d[0]]}")
    # predictions
    for x in kmers[:2000]:
        if ref.predict(int(x)) != port.predict(int(x)):
            bad.append(f"predict({int(x)})")
            break
    # variable-length strings through kmerizeAdjusted
    nd = 0
    for s, L in var_len_queries(genome, k, 300, rng):
        xa = ref.kmerize_adjusted(L, s)
        if xa != O.kmerize_adjusted(k, L, s):
            bad.append("kmerizeAdjusted")
            break
        if port.predict(xa) >= n:
            continue
        a, b = ref.query_str(s, xa), port.query_str(s, xa)
        nd += a != b
    if nd:
        bad.append(f"variable-length plQuery: {nd} diffs")
    frac_neg1 = float(np.mean(pr == -1))
    print(f"{tag:32s} n={n:<9d} nb={ref.nb:<2d} five={ref.five} perfect={ref.perfect} "
          f"queries={int(keep.sum())} oob_excluded={int((~keep).sum())} probes/q={probes / len(kmers):.2f} "
          f"-1={frac_neg1:.3f} -> {'OK' if not bad else 'MISMATCH ' + '; '.join(bad)}", flush=True)
    ref.close()
    port.close()
    return not bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    O.build()
    ok = True
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        for name, g in fixture_genomes(args.big):
            n = len(g)
            for k in (11, 16, 21, 31):
                if n < 4 * k:
                    continue
                for nb in (-1, 4, 8, 12):
                    if nb > 2 * k:
                        continue
                    ok &= run_case(name, g, k, nb, tmp)
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

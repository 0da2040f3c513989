#!/usr/bin/env python
"""Generate tests/golden/chr3_10M.npz from the UNMODIFIED reference: the real-genome fixture of SURVEY section 4 / 8c,
Complete-Striped-Smith-Waterman-Library/demo/10M.fa (human chr3:50000-10050000; 9,850,001 bases after the reference's
N-stripping).  Real sequence is the hostile case for this path: repeats give error bounds in the thousands
(maxOver = 4231, maxUnder = 4498 against 25 / 12 on a random genome of 300 times the size), 6.8 % of its 21-mers occur more
than once (answers depend on the probe sequence), and the prefix deltas of the rank lines overflow far more often.

The fixture holds the cleaned genome 2-bit packed (2.5 MB; data shipped with the reference, not source code), what the
reference's own constructor printed for it, and the reference's plQuery answers for 100 000 present and 100 000 mutated
21-mers (regenerated from the genome by the seeded generators of tests/_oracle.py).  Runs only where /root/reference exists:

    python tests/golden/make_golden_chr3.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

FA = "/root/reference/Complete-Striped-Smith-Waterman-Library/demo/10M.fa"
K = 21
NQ = 100_000


def pack2(genome: bytes) -> np.ndarray:
    code = np.zeros(256, dtype=np.uint8)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    c = code[np.frombuffer(genome, dtype=np.uint8)]
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]


def unpack2(packed: np.ndarray, n: int) -> bytes:
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    c = np.stack([(packed >> 6) & 3, (packed >> 4) & 3, (packed >> 2) & 3, packed & 3], axis=1).reshape(-1)[:n]
    return letters[c].tobytes()


def main():
    O.build()
    g, ends = O.clean_fasta_text(open(FA, "rb").read())
    n = len(g)
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        fa = os.path.join(tmp, "chr3.fa")
        O.write_fasta(fa, g)
        # the suffix array comes from the reference's own libdivsufsort (suffixarray/refToSuffixArray.sh); the .sa file's
        # second half (LCP, suffixarray/addlcp.cpp) is written by the port, which the small fixtures pin byte for byte
        sa = O.divsufsort(g).astype(np.uint32)
        port = O.Port.from_memory(g, sa=sa, k=K)
        port.write_sa(fa + ".sa")
        ref = O.Ref(fa, fa + ".sa", fa + ".sap", k=K)  # the reference builds its model itself (sapling_api.h:384-487)
        assert (ref.n, ref.nb) == (n, port.nb) and ref.five == port.five and ref.perfect == port.perfect
        present, pos = O.present_queries(g, K, NQ)
        kmers = np.concatenate([present, O.mutate_queries(present, K, every=1)])
        ok = np.array([port.predict(int(x)) < n for x in kmers], dtype=bool)  # reference UB otherwise
        assert ok.all(), "the fixture's queries are regenerated in the tests: none may be excluded"
        answers = ref.query_batch(kmers, nthreads=8)
        exp_port, probes, _ = port.query_batch(kmers, nthreads=8, stats=True)
        assert np.array_equal(exp_port, answers), "oracle port differs from the reference on chr3"
        out = os.path.join(HERE, "chr3_10M.npz")
        np.savez_compressed(out, packed=pack2(g), n=np.int64(n), k=np.int32(K), nb=np.int32(ref.nb),
                            five=np.array(ref.five, dtype=np.int32), perfect=np.int64(ref.perfect),
                            answers=answers.astype(np.int32), excluded_ub=np.int64((~ok).sum()),
                            probes_per_query=np.float64(probes / len(kmers)),
                            sa_crc=np.int64(int(np.bitwise_xor.reduce(sa.astype(np.uint64) * np.arange(1, n + 1, dtype=np.uint64)))))
        print(f"wrote {out}: n={n} nb={ref.nb} five={ref.five} perfect={ref.perfect} queries={len(kmers)} "
              f"(-1: {(answers == -1).sum()}, excluded UB: {(~ok).sum()}) probes/query={probes / len(kmers):.2f} "
              f"size={os.path.getsize(out) / 1e6:.2f} MB")
        ref.close()
        port.close()


if __name__ == "__main__":
    main()

"""Deterministic fixture genomes and query sets shared by the CPU and GPU parity tests."""
import numpy as np

import _oracle as O


def _freq(rng, n, w):
    w = np.array(w, dtype=float)
    return bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=n, p=w / w.sum())])


def small_genomes():
    """name -> genome bytes.  Random, GC-skewed and tandem-repeat recipes restated from
    eval/SuffixArraySim/SuffixArraySimulatedSequences.java:13-26 (reference), sized for seconds."""
    rng = np.random.default_rng(20240607)
    g = {}
    g["rand20k"] = O.synth_genome(O.SEED_G, 20000)
    g["rand200k"] = O.synth_genome(O.SEED_G + 7, 200000)
    g["gc2332"] = _freq(rng, 40000, [2, 3, 3, 2])
    g["gc1991"] = _freq(rng, 40000, [1, 9, 9, 1])
    g["gc0110"] = _freq(rng, 20000, [0, 1, 1, 0])
    g["polyC"] = b"C" * 3000
    for unit in ("CT", "CAT", "ACGT", "ACTTCA", "ACGTCGTAGTACTACG"):
        g["tandem_" + unit] = (unit * (12000 // len(unit) + 1)).encode()[:12000]
    g["tandem50"] = (_freq(rng, 50, [1, 1, 1, 1]) * 300)[:12000]
    # a genome with a long exact repeat and a poly-A tail (end-of-text handling)
    core = O.synth_genome(O.SEED_G + 99, 30000)
    g["repeat_tailA"] = core + core[5000:15000] + b"A" * 100
    return g


_CODE = np.zeros(256, dtype=np.uint64)
_CODE[ord("C")] = 1
_CODE[ord("G")] = 2
_CODE[ord("T")] = 3


def pack_genome(genome: bytes, pad_words=4):
    """2-bit pack, 32 bases per uint64, base i in the top bits first (sapling_b200/csrc/common.cuh)."""
    n = len(genome)
    nw = (n + 31) // 32 + pad_words
    codes = np.zeros(nw * 32, dtype=np.uint64)
    codes[:n] = _CODE[np.frombuffer(genome, dtype=np.uint8)]
    codes = codes.reshape(nw, 32)
    shifts = (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))
    return np.bitwise_or.reduce(codes << shifts, axis=1)


def pack_strings(strings):
    """(words, word_off) in the layout sapling_b200_query_str_batch builds."""
    offs, words = [], []
    total = 0
    for s in strings:
        offs.append(total)
        nw = (len(s) + 31) // 32 + 1
        w = [0] * nw
        for j, c in enumerate(s):
            v = {67: 1, 71: 2, 84: 3}.get(c, 0)
            w[j >> 5] |= v << (62 - 2 * (j & 31))
        words.extend(w)
        total += nw
    return np.array(words, dtype=np.uint64), np.array(offs, dtype=np.uint64)


def query_mix(genome: bytes, k, nq, seed=0):
    """present + mutated + uniformly random k-mers"""
    present, _ = O.present_queries(genome, k, nq, seed=O.SEED_Q + seed)
    mutated = O.mutate_queries(present, k, seed=O.SEED_M + seed)
    rng = np.random.default_rng(1000 + seed)
    hi = 1 << (2 * k)
    randq = rng.integers(0, hi, size=max(1, nq // 4), dtype=np.uint64) if hi <= (1 << 63) else \
        rng.integers(0, 1 << 63, size=max(1, nq // 4), dtype=np.uint64)
    return np.concatenate([present, mutated, randq])


def var_len_strings(genome: bytes, k, count, seed=0):
    """strings of the sapling_example sweep lengths k-10..k+80 (sapling_example.cpp:93-98), half of them
    mutated at one position"""
    rng = np.random.default_rng(77 + seed)
    n = len(genome)
    out = []
    for L in (max(1, k - 10), k, k + 10, k + 20, k + 30, k + 80):
        if L >= n:
            continue
        for p in rng.integers(0, n - L, size=count):
            s = bytearray(genome[p:p + L])
            if rng.random() < 0.5:
                j = int(rng.integers(0, L))
                s[j] = b"ACGT"[(b"ACGT".index(s[j]) + 1 + int(rng.integers(0, 3))) % 4]
            out.append(bytes(s))
    return out


def narrow_model(xlist, ylist, k, nb):
    """The 8-byte-per-bucket device layout of sapling_b200/csrc/model.cu, restated in numpy:
    returns (uint32 array of shape (B, 2) = {xoff | fill flag, y}, ok)."""
    shift = 2 * k - nb
    B = 1 << nb
    if shift < 0 or shift > 31:
        return None, False
    x = np.asarray(xlist[:B], dtype=np.int64)
    y = np.asarray(ylist[:B], dtype=np.int64)
    base = np.arange(B, dtype=np.int64) << shift
    inside = (x >= base) & (x < base + (1 << shift))
    src = np.where(inside, np.arange(B), x >> shift).astype(np.int64)
    ok = bool(((x >= 0) & (y >= 0) & (y <= 0xFFFFFFFF)).all() and (x[src] == x).all() and (y[src] == y).all()
              and (inside | (x < base)).all())
    xoff = np.where(inside, x - base, 0x80000000 | (np.arange(B) - src)).astype(np.uint32)
    table = np.stack([xoff, y.astype(np.uint32)], axis=1)
    pad = np.array([[0x80000000, 0]], dtype=np.uint32)  # entry B: flagged, lets the kernels read bucket b + 1 unconditionally
    return np.ascontiguousarray(np.concatenate([table, pad])), ok

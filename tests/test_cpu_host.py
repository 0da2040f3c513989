"""CPU-side tests (-m "not gpu"): the oracle against the golden vectors, the host logic of the
library that needs no GPU, the C-ABI export list, and the device query code compiled for the host
(tests/sim) against the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _fixtures as F
import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    G.build()
    return True


def test_capi_exports_every_declared_symbol(built):
    import sapling_b200
    hdr = open(os.path.join(ROOT, "include", "sapling_b200.h")).read()
    declared = set(re.findall(r"\b(sapling_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = C.CDLL(sapling_b200.lib_path())
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/sapling_b200.h but not exported"
    # the Python binding covers the same list
    from sapling_b200.api import SYMBOLS
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)


def test_no_cpu_fallback(built):
    """Without a GPU every constructor must fail loudly (never silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sapling_b200 as S
    with pytest.raises(S.SaplingError, match="no usable CUDA device|not a Blackwell"):
        S.Sapling.from_memory(b"ACGT" * 100, k=11)
    with pytest.raises(S.SaplingError):
        S.Sapling.synthetic(1, 10000, k=11)


def test_product_does_not_touch_oracle():
    """Nothing under sapling_b200/ or include/ may reference oracle/ (the checker is not the product)."""
    for base in ("sapling_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "oracle/" not in txt and "sapling_oracle" not in txt and "_oracle" not in txt, (dp, fn)


def test_kmerize_host(built, oracle_built):
    import sapling_b200 as S
    rng = np.random.default_rng(3)
    for k in (1, 11, 16, 21, 31):
        for _ in range(50):
            L = int(rng.integers(1, 120))
            s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=max(L, k)))
            assert S.kmerize(k, s) == O.kmerize(k, s)
            assert S.kmerize_adjusted(k, L, s) == O.kmerize_adjusted(k, L, s)


_SIM = None


def _sim():
    """The device query code compiled for the host (tests/sim/sim_query.cpp), built once per session and only when a
    source is newer than the library; written under a temporary name first so that parallel test workers never load a
    half-written file."""
    global _SIM
    if _SIM is not None:
        return _SIM
    so = os.path.join(ROOT, "build", "libsim_query.so")
    src = os.path.join(ROOT, "tests", "sim", "sim_query.cpp")
    deps = [src] + [os.path.join(ROOT, "sapling_b200", "csrc", f) for f in ("query.cuh", "common.cuh")]
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        tmp = f"{so}.{os.getpid()}.tmp"
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                        "-I/usr/local/cuda/include", "-o", tmp, src], check=True)
        os.replace(tmp, so)
    L = C.CDLL(so)
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.sim_kmer_batch.argtypes = [O.u64p, O.u32p, O.i64p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int, O.u64p,
                                 C.c_size_t, O.i64p, C.POINTER(C.c_uint64), C.c_void_p, O.i64p]
    L.sim_string_batch.argtypes = [O.u64p, O.u32p, O.i64p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int, O.u64p,
                                   O.u64p, O.u32p, O.u32p, O.i64p, C.c_size_t, O.i64p, C.POINTER(C.c_uint64),
                                   C.c_void_p, O.i64p]
    L.sim_kmer_batch_packed.argtypes = [O.u64p, O.u32p, O.i64p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int, O.u64p,
                                        C.c_size_t, O.i64p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_void_p,
                                        C.POINTER(C.c_uint64)]
    L.sim_kmer_batch_lean.restype = C.c_int
    L.sim_kmer_batch_lean.argtypes = [O.u64p, O.u32p, O.i64p, C.c_uint64, C.c_int, C.c_int, i32p, C.c_int, O.u64p,
                                      C.c_size_t, O.i64p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int]
    _SIM = L
    return L


@pytest.mark.parametrize("name", ["rand20k", "gc1991", "gc0110", "polyC", "tandem_CT", "tandem50", "repeat_tailA"])
def test_lean_kmer_replay_on_host_matches_oracle(oracle_built, name):
    """query.cuh kmer_replay32 (32-bit ranks, the nine cases of the replay folded into one update; what the batch
    kernels run) compiled for the host == oracle, on all three layouts: suffix-array sector + packed genome, inline
    prefixes, rank lines (overlapping / tiling, prefixes shorter than k, escapes); the model's own error bounds, bounds
    that collapse the left window to rank 0 (long-window shortcut), tiny bounds; compat and 64-bit-safe
    window arithmetic."""
    L = _sim()
    g = F.small_genomes()[name]
    n = len(g)
    for k, nb in ((21, -1), (11, 4), (31, 10), (16, -1), (32, 12)):
        if n < 4 * k:
            continue
        if k == 32:
            continue  # oracle-undefined
        base = O.Port.from_memory(g, nb=nb, k=k)
        packed, sa = F.pack_genome(g), base.sa
        model = np.ascontiguousarray(np.stack([base.xlist, base.ylist], axis=1).reshape(-1))
        kmers = F.query_mix(g, k, 3000)
        tail = np.array([O.kmerize(k, g[i:i + k] + b"A" * k) for i in range(max(0, n - 40), n)], dtype=np.uint64)
        kmers = np.concatenate([kmers, tail])
        f0 = list(base.five)
        for five_t in (f0, [f0[0], f0[1], f0[2], f0[3], 1 << 30], [f0[0], 2, f0[2], 1, 1 << 30], [2, 2, 1, 1, 1],
                       [1, 40, 1, 9, 30]):
            port = O.Port.from_parts(g, sa, k, base.nb, base.xlist, base.ylist, five_t)
            five = np.array(five_t, dtype=np.int32)
            exp, _, oob = port.query_batch(kmers, nthreads=2, stats=True)
            cases = [(0, 0, 3), (2, 6, 3), (2, 12, 4), (2, 21, 3), (2, 32, 4), (2, 16, 4),
                     (3, 6, 4), (3, 12, 3), (3, 21, 4), (3, 32, 3), (3, 16, 3),
                     (4, 6, 4), (4, 12, 4), (4, 21, 4), (4, 32, 4), (4, 16, 4), (4, 24, 4),
                     (5, 6, 4), (5, 12, 3), (5, 21, 4), (5, 32, 3), (5, 24, 4)]
            cases += [(1, b, 3) for b in (27, 32) if k <= b]
            for mode, bases, shift in cases:
                out = np.empty(len(kmers), dtype=np.int64)
                c = C.c_uint64(0)
                rc = L.sim_kmer_batch_lean(packed, sa, model, n, k, base.nb, five, 1, kmers, len(kmers), out, C.byref(c),
                                           mode, bases, shift)
                assert rc == 0
                bad = np.nonzero(out != exp)[0]
                assert len(bad) == 0 and c.value == oob, (name, k, nb, five_t, mode, bases, shift, bad[:5], out[bad[:5]],
                                                          exp[bad[:5]])
            # 64-bit-safe window arithmetic (SAPLING_B200_NO_COMPAT) has no oracle: lean replay == general replay
            gen = np.empty(len(kmers), dtype=np.int64)
            last = np.array([base.xlist[-1], base.ylist[-1]], dtype=np.int64)
            L.sim_kmer_batch(packed, sa, model, n, k, base.nb, five, 0, kmers, len(kmers), gen, C.byref(c), None, last)
            for mode, bases, shift in cases[:3] + [(4, 12, 4), (4, 21, 4)]:
                out = np.empty(len(kmers), dtype=np.int64)
                assert L.sim_kmer_batch_lean(packed, sa, model, n, k, base.nb, five, 0, kmers, len(kmers), out,
                                             C.byref(c), mode, bases, shift) == 0
                assert np.array_equal(out, gen), (name, k, nb, five_t, mode, "no-compat")
            port.close()
        base.close()


@pytest.mark.parametrize("name", ["rand20k", "gc1991", "gc0110", "polyC", "tandem_CT", "tandem50", "repeat_tailA"])
def test_rank_line_query_code_on_host_matches_oracle(oracle_built, name):
    """The rank-line layout (common.cuh pack_rank_sector + query.cuh SaPacked, the default for genomes >= 50 Mbp)
    compiled for the host == oracle: overlapping and tiling lines, prefixes shorter than / as long as / longer than k
    (genome fallback on ties), prefixes wide enough that 21-bit deltas overflow (escapes), suffixes at the end of the
    text, and the collapsed left window of SURVEY F5."""
    L = _sim()
    g = F.small_genomes()[name]
    n = len(g)
    escapes = 0
    for k, nb in ((21, -1), (11, 4), (31, 10), (16, -1)):
        if n < 4 * k:
            continue
        base = O.Port.from_memory(g, nb=nb, k=k)
        packed, sa = F.pack_genome(g), base.sa
        model = np.ascontiguousarray(np.stack([base.xlist, base.ylist], axis=1).reshape(-1))
        kmers = F.query_mix(g, k, 3000)
        # every suffix of the last 40 positions as a query too (short suffixes are escaped entries)
        tail = np.array([O.kmerize(k, g[i:i + k] + b"A" * k) for i in range(max(0, n - 40), n)], dtype=np.uint64)
        kmers = np.concatenate([kmers, tail])
        f0 = list(base.five)
        for five_t in (f0, [f0[0], f0[1], f0[2], f0[3], 1 << 30], [f0[0], 2, f0[2], 1, 1 << 30]):
            port = O.Port.from_parts(g, sa, k, base.nb, base.xlist, base.ylist, five_t)
            five = np.array(five_t, dtype=np.int32)
            exp, _, oob = port.query_batch(kmers, nthreads=2, stats=True)
            for bases in (6, 12, 16, 21, 32):
                for shift in (3, 4):
                    out = np.empty(len(kmers), dtype=np.int64)
                    c, e = C.c_uint64(0), C.c_uint64(0)
                    L.sim_kmer_batch_packed(packed, sa, model, n, k, base.nb, five, 1, kmers, len(kmers), out, C.byref(c),
                                            bases, shift, None, C.byref(e))
                    assert np.array_equal(out, exp) and c.value == oob, (name, k, nb, five_t, bases, shift)
                    escapes += e.value
            port.close()
        base.close()
    assert escapes > 0  # at least the short suffixes at the end of the text


@pytest.mark.parametrize("name", ["rand20k", "gc1991", "gc0110", "polyC", "tandem_CT", "tandem50", "repeat_tailA"])
def test_device_query_code_on_host_matches_oracle(oracle_built, name):
    """query.cuh (the device replay of plQuery) compiled for the host == oracle, k-mers and strings."""
    L = _sim()
    g = F.small_genomes()[name]
    n = len(g)
    for k, nb in ((21, -1), (11, 4), (31, 10), (16, -1)):
        if n < 4 * k:
            continue
        base = O.Port.from_memory(g, nb=nb, k=k)
        packed, sa = F.pack_genome(g), base.sa
        model = np.ascontiguousarray(np.stack([base.xlist, base.ylist], axis=1).reshape(-1))
        last = np.array([base.xlist[-1], base.ylist[-1]], dtype=np.int64)
        narrow, nok = F.narrow_model(base.xlist, base.ylist, k, base.nb)
        layouts = [None] + ([narrow.ctypes.data_as(C.c_void_p)] if nok else [])
        assert nok or 2 * k - base.nb > 31
        kmers = F.query_mix(g, k, 3000)
        strs = F.var_len_strings(g, k, 25)
        words, offs = F.pack_strings(strs)
        slens = np.array([len(s) for s in strs], dtype=np.uint32)
        km = np.array([O.kmerize_adjusted(k, len(s), s) for s in strs], dtype=np.int64)
        # the model's own error bounds, then bounds that collapse the left window to rank 0 the way the reference's
        # (int)predicted cast does beyond 2^31: exercises the long-window shortcut of pl_query_from
        f0 = list(base.five)
        for five_t in (f0, [f0[0], f0[1], f0[2], f0[3], 1 << 30], [f0[0], 2, f0[2], 1, 1 << 30]):
            port = O.Port.from_parts(g, sa, k, base.nb, base.xlist, base.ylist, five_t)
            five = np.array(five_t, dtype=np.int32)
            exp, _, oob = port.query_batch(kmers, nthreads=2, stats=True)
            for nptr in layouts:  # wide table, then the narrow 8-byte layout
                out = np.empty(len(kmers), dtype=np.int64)
                c = C.c_uint64(0)
                L.sim_kmer_batch(packed, sa, model, n, k, base.nb, five, 1, kmers, len(kmers), out, C.byref(c), nptr, last)
                assert np.array_equal(out, exp) and c.value == oob, (name, k, nb, five_t)
            exp2 = np.array([port.query_str(s, x) for s, x in zip(strs, km)], dtype=np.int64)
            out2 = np.empty(len(strs), dtype=np.int64)
            L.sim_string_batch(packed, sa, model, n, k, base.nb, five, 1, words, offs, slens, slens, km, len(strs), out2,
                               C.byref(c), layouts[-1], last)
            assert np.array_equal(out2, exp2), (name, k, nb, five_t)
            port.close()
        base.close()
    tried, ok = C.c_uint64(0), C.c_uint64(0)
    L.sim_skip_counters(C.byref(tried), C.byref(ok))
    assert tried.value > 1000 and ok.value > 1000, (tried.value, ok.value)  # the shortcut was exercised (cumulative)
    print(f"long-window shortcut: tried {tried.value}, accepted {ok.value}")


@pytest.mark.parametrize("seed", range(40))
def test_randomised_replay_against_oracle(oracle_built, seed):
    """Randomised differential test of the device replay code (host simulation) against the oracle: random genome
    recipes (uniform, skewed, short tandem units, planted repeats), random k, bucket count and error bounds (bounds
    smaller than the model's true errors included: the replay must reproduce the reference's wrong answers too), every
    replay variant and layout the batch kernels can run."""
    L = _sim()
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(400, 6000))
    kind = seed % 4
    if kind == 0:
        g = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])
    elif kind == 1:
        w = rng.random(4) ** 3 + 1e-3
        g = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=n, p=w / w.sum())])
    elif kind == 2:
        unit = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 9)))])
        g = (unit * (n // len(unit) + 1))[:n]
    else:
        core = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n // 2)])
        g = core + core[: n // 3] + b"A" * int(rng.integers(0, 40)) + core[n // 4:]
        n = len(g)
    k = int(rng.integers(6, 32))
    nb = int(rng.integers(1, min(2 * k, 14) + 1))
    base = O.Port.from_memory(g, nb=nb, k=k)
    packed, sa = F.pack_genome(g), base.sa
    model = np.ascontiguousarray(np.stack([base.xlist, base.ylist], axis=1).reshape(-1))
    last = np.array([base.xlist[-1], base.ylist[-1]], dtype=np.int64)
    narrow, nok = F.narrow_model(base.xlist, base.ylist, k, base.nb)
    kmers = F.query_mix(g, k, 1500, seed=seed)
    f0 = list(base.five)
    bounds = [f0, [int(rng.integers(0, 6)), int(rng.integers(0, 6)), 1, int(rng.integers(0, 4)), int(rng.integers(0, 4))],
              [f0[0], f0[1], f0[2], f0[3], 1 << 30]]
    for five_t in bounds:
        port = O.Port.from_parts(g, sa, k, base.nb, base.xlist, base.ylist, five_t)
        five = np.array(five_t, dtype=np.int32)
        exp, _, oob = port.query_batch(kmers, nthreads=2, stats=True)
        port.close()
        for nptr in [None] + ([narrow.ctypes.data_as(C.c_void_p)] if nok else []):
            out = np.empty(len(kmers), dtype=np.int64)
            c = C.c_uint64(0)
            L.sim_kmer_batch(packed, sa, model, n, k, base.nb, five, 1, kmers, len(kmers), out, C.byref(c), nptr, last)
            assert np.array_equal(out, exp) and c.value == oob, (seed, n, k, nb, five_t, "general", nptr is not None)
        pb = int(rng.integers(4, 33))
        cases = [(0, 0, 3), (2, pb, 3), (2, pb, 4), (3, pb, 3), (3, pb, 4), (4, pb, 4), (5, pb, 4), (5, pb, 3)]
        cases += [(1, b, 3) for b in (27, 32) if k <= b]
        for mode, bases, shift in cases:
            out = np.empty(len(kmers), dtype=np.int64)
            c = C.c_uint64(0)
            rc = L.sim_kmer_batch_lean(packed, sa, model, n, k, base.nb, five, 1, kmers, len(kmers), out, C.byref(c),
                                       mode, bases, shift)
            assert rc == 0
            bad = np.nonzero(out != exp)[0]
            assert len(bad) == 0 and c.value == oob, (seed, n, k, nb, five_t, mode, bases, shift, bad[:5], out[bad[:5]],
                                                      exp[bad[:5]])
    base.close()


@pytest.mark.parametrize("seed", range(10))
def test_randomised_string_queries_against_oracle(oracle_built, seed):
    """The string form of the query (plQuery(s, kmer, length) with length != k: kmerizeAdjusted, the gallop loops
    sapling_api.h:184-196,:229-241) on random genomes / k / nb, wide and narrow model layouts, against the oracle."""
    L = _sim()
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(800, 5000))
    if seed % 3 == 0:
        g = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])
    elif seed % 3 == 1:
        unit = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 9)))])
        g = (unit * (n // len(unit) + 1))[:n]
    else:
        core = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n // 2)])
        g = core + core[: n // 3] + b"A" * int(rng.integers(0, 40))
        n = len(g)
    k = int(rng.integers(11, 32))
    nb = int(rng.integers(1, min(2 * k, 12) + 1))
    base = O.Port.from_memory(g, nb=nb, k=k)
    packed, sa = F.pack_genome(g), base.sa
    model = np.ascontiguousarray(np.stack([base.xlist, base.ylist], axis=1).reshape(-1))
    last = np.array([base.xlist[-1], base.ylist[-1]], dtype=np.int64)
    narrow, nok = F.narrow_model(base.xlist, base.ylist, k, base.nb)
    strs = F.var_len_strings(g, k, 20, seed=seed)
    words, offs = F.pack_strings(strs)
    slens = np.array([len(s) for s in strs], dtype=np.uint32)
    km = np.array([O.kmerize_adjusted(k, len(s), s) for s in strs], dtype=np.int64)
    f0 = list(base.five)
    for five_t in (f0, [f0[0], f0[1], f0[2], f0[3], 1 << 30]):
        port = O.Port.from_parts(g, sa, k, base.nb, base.xlist, base.ylist, five_t)
        five = np.array(five_t, dtype=np.int32)
        exp = np.array([port.query_str(s, x) for s, x in zip(strs, km)], dtype=np.int64)
        port.close()
        for nptr in [None] + ([narrow.ctypes.data_as(C.c_void_p)] if nok else []):
            out = np.empty(len(strs), dtype=np.int64)
            c = C.c_uint64(0)
            L.sim_string_batch(packed, sa, model, n, k, base.nb, five, 1, words, offs, slens, slens, km, len(strs), out,
                               C.byref(c), nptr, last)
            assert np.array_equal(out, exp), (seed, n, k, nb, five_t, nptr is not None)
    base.close()


def test_seed_batch_port_matches_reference_methods(oracle_built, tmp_path):
    """The oracle's restatement of the align.cpp seed loop == the same loop driven through the unmodified reference's
    own kmerize / plQuery / countHitsLeft/Right (oracle/ref_harness.cpp).  Needs /root/reference (skipped on the GPU box,
    where the committed golden fixture tests/golden/seeds_*.npz pins it instead)."""
    if not O.ref_available():
        pytest.skip("reference not built here")
    for name, k in (("rand20k", 16), ("gc1991", 16), ("tandem50", 16), ("repeat_tailA", 21)):
        g = F.small_genomes()[name]
        reads, _ = O.simulate_reads(g, 120, min(150, len(g) // 4))
        reads += [b"ACGT", g[100:100 + k], b"N" * 60, g[-150:], g[:150]]
        fa = tmp_path / f"{name}.fa"
        O.write_fasta(str(fa), g)
        ref = O.Ref(str(fa), str(tmp_path / f"{name}.sa"), str(tmp_path / f"{name}.{k}.sap"), k=k)
        port = O.Port.from_memory(g, k=k)
        for num_seeds, max_hits in ((7, 32), (2, 5)):
            a, b = ref.seed_batch(reads, num_seeds, max_hits), port.seed_batch(reads, num_seeds, max_hits)
            # countHitsLeft on the LAST rank reads lcp[n-1], one past the reference's n-1 entries (sapling_api.h:287):
            # undefined there; port and GPU define that flag as 0.  Excluded from the comparison, everything else equal.
            defined = ~((a[0] >= 0) & (a[1] == len(g) - 1))
            for x, y in zip(a, b):
                assert np.array_equal(x[defined], y[defined]), (name, k, num_seeds)
        ref.close()
        port.close()

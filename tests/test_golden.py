"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

CPU (-m "not gpu"): the oracle port reproduces every golden fixture (this is what pins the oracle on a
box without /root/reference).  GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

import _oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                  if not os.path.basename(p).startswith(("seeds.", "chr3_")))
SEED_FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "seeds.*.npz")))
SEED_IDS = [os.path.basename(p)[:-4] for p in SEED_FIXTURES]
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def _load(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["genome"] = d["genome"].tobytes()
    blob = d["strings"].tobytes()
    offs = np.concatenate([[0], np.cumsum(d["string_lens"])]).astype(np.int64)
    d["strs"] = [blob[offs[i]:offs[i + 1]] for i in range(len(d["string_lens"]))]
    return d


def test_fixtures_present():
    assert len(FIXTURES) >= 10


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_oracle_port_reproduces_reference(oracle_built, path):
    d = _load(path)
    port = O.Port.from_memory(d["genome"], nb=int(d["nb_arg"]), k=int(d["k"]))
    assert np.array_equal(port.sa, d["sa"])
    assert port.nb == int(d["nb"])
    assert np.array_equal(port.xlist, d["xlist"]) and np.array_equal(port.ylist, d["ylist"])
    assert port.five == tuple(int(v) for v in d["five"])
    assert port.perfect == int(d["perfect"])
    assert np.array_equal(port.query_batch(d["kmers"]), d["answers"])
    got = [port.query_str(s, int(x)) for s, x in zip(d["strs"], d["string_kmers"])]
    assert got == [int(v) for v in d["string_answers"]]
    port.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_reproduces_reference(path):
    import sapling_b200 as S
    d = _load(path)
    k = int(d["k"])
    ix = S.Sapling.from_memory(d["genome"], None, numBuckets=int(d["nb_arg"]), k=k)
    assert np.array_equal(ix.rev(), d["sa"])
    assert ix.buckets == int(d["nb"])
    x, y = ix.model()
    assert np.array_equal(x, d["xlist"]) and np.array_equal(y, d["ylist"])
    assert ix.five == tuple(int(v) for v in d["five"])
    assert ix.perfectPredictions == int(d["perfect"])
    assert np.array_equal(ix.queryBatch(d["kmers"]), d["answers"])
    assert np.array_equal(ix.plQueryBatch(d["strs"], d["string_kmers"]), d["string_answers"])
    for s, x, e in list(zip(d["strs"], d["string_kmers"], d["string_answers"]))[:8]:
        assert ix.plQuery(s, int(x), len(s)) == int(e)
    ix.close()


def _load_seeds(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["genome"] = d["genome"].tobytes()
    blob = d["reads"].tobytes()
    offs = np.concatenate([[0], np.cumsum(d["read_lens"])]).astype(np.int64)
    d["read_list"] = [blob[offs[i]:offs[i + 1]] for i in range(len(d["read_lens"]))]
    return d


def _check_seeds(d, got):
    ok = d["defined"]
    for a, what in zip(got, ("ref_pos", "sa_pos", "left", "right")):
        assert np.array_equal(a[ok], d[what][ok]), what
    assert (d["ref_pos"] >= 0).sum() > 50


@pytest.mark.parametrize("path", SEED_FIXTURES, ids=SEED_IDS)
def test_oracle_port_reproduces_reference_seed_loop(oracle_built, path):
    """align.cpp:259-300 through the reference's own methods (golden) == the oracle's restatement."""
    d = _load_seeds(path)
    port = O.Port.from_memory(d["genome"], k=int(d["k"]))
    _check_seeds(d, port.seed_batch(d["read_list"], int(d["num_seeds"]), int(d["max_hits"])))
    port.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", SEED_FIXTURES, ids=SEED_IDS)
def test_cuda_reproduces_reference_seed_loop(path):
    import sapling_b200 as S
    d = _load_seeds(path)
    ix = S.Sapling.from_memory(d["genome"], None, k=int(d["k"]), flags=S.QUIET | S.KEEP_BUILD)
    _check_seeds(d, ix.seedBatch(d["read_list"], int(d["num_seeds"]), int(d["max_hits"])))
    ix.close()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_reference_struct_filled_from_parts_equals_constructor(tmp_path):
    """oracle/ref_harness.cpp ref_from_parts (the reference's struct filled member by member, used for the 3.1 Gbp
    parity and CPU-baseline legs) answers exactly like the reference built through its own constructor and files."""
    g = O.synth_genome(O.SEED_G + 3, 150000)
    fa = str(tmp_path / "g.fa")
    O.write_fasta(fa, g)
    for k, nb in ((21, -1), (16, 9), (31, 12)):
        ref = O.Ref(fa, fa + f".{k}.sa", fa + f".{k}.{nb}.sap", nb=nb, k=k)
        parts = O.Ref.from_parts(ref.genome, ref.sa.astype(np.uint32), k, ref.nb, ref.xlist, ref.ylist, ref.five,
                                 nthreads=2, chunk=40000)
        km, _ = O.present_queries(g, k, 30000)
        km = np.concatenate([km, O.mutate_queries(km, k)])
        assert np.array_equal(parts.query_batch(km, nthreads=2), ref.query_batch(km, nthreads=2))
        assert (parts.n, parts.k, parts.nb, parts.five) == (ref.n, ref.k, ref.nb, ref.five)
        parts.close()
        ref.close()


@pytest.mark.skipif(not O.divsuf_available(), reason="oracle/_ref/libdivsuf_ref.so not built (needs /root/reference at build time)")
def test_port_suffix_array_and_equal_range_pinned_to_libdivsufsort():
    """The reference's own cross-checks: the oracle port's suffix array == divsufsort()'s (and passes its
    sufcheck), and the port's match-range helper so_equal_range == sa_search (lib/utils.c:259-326) for present and absent
    patterns of every k the tests use, 32 included (where the reference's Sapling itself is undefined)."""
    import _fixtures as F
    for name in ("rand20k", "gc1991", "tandem50", "repeat_tailA", "polyC"):
        g = F.small_genomes()[name]
        port = O.Port.from_memory(g, k=min(21, len(g) // 8))
        sa = port.sa
        assert np.array_equal(O.divsufsort(g), sa.astype(np.int64)), name
        assert O.sufcheck(g, sa) == 0
        for k in (11, 16, 21, 31, 32):
            if len(g) < 4 * k:
                continue
            km, _ = O.present_queries(g, k, 600)
            km = np.concatenate([km, O.mutate_queries(km, k, every=1)])
            left, cnt = O.sa_search_batch(g, sa, km, k)
            for i in range(len(km)):
                lb, ub = port.equal_range(O.unpack_kmer(int(km[i]), k))
                assert ub - lb == cnt[i] and (cnt[i] == 0 or lb == left[i]), (name, k, i)
            assert (cnt[:600] > 0).all()
        port.close()


CHR3 = os.path.join(HERE, "golden", "chr3_10M.npz")


def load_chr3():
    """The real-genome fixture (tests/golden/make_golden_chr3.py): genome bytes, expected stats, queries, the reference's
    answers."""
    sys_path = os.path.join(HERE, "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_chr3", os.path.join(sys_path, "make_golden_chr3.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(CHR3)
    n, k = int(z["n"]), int(z["k"])
    g = mod.unpack2(z["packed"], n)
    present, _ = O.present_queries(g, k, mod.NQ)
    kmers = np.concatenate([present, O.mutate_queries(present, k, every=1)])
    return g, k, int(z["nb"]), tuple(int(v) for v in z["five"]), int(z["perfect"]), kmers, z["answers"].astype(np.int64)


def test_port_reproduces_reference_on_real_chr3():
    """human chr3 10 Mbp (the SSW demo data shipped with the reference) -- n = 9,850,001, nb = 19,
    maxOver = 4231, maxUnder = 4498, mostOver = 29, mostUnder = 28, meanError = 11, perfect = 1,635,411 -- and the
    reference's answers to 200 000 present / mutated 21-mers, reproduced by the oracle port from the genome alone."""
    g, k, nb, five, perfect, kmers, answers = load_chr3()
    assert (len(g), nb, five, perfect) == (9_850_001, 19, (4231, 4498, 11, 29, 28), 1_635_411)
    port = O.Port.from_memory(g, k=k)
    assert (port.nb, port.five, port.perfect) == (nb, five, perfect)
    assert np.array_equal(port.query_batch(kmers, nthreads=4), answers)
    port.close()


@pytest.mark.gpu
def test_cuda_reproduces_reference_on_real_chr3(monkeypatch):
    """The same fixture through the CUDA path: suffix array + model built on the GPU give the reference's statistics, and
    the partitioned and unpartitioned kernels give the reference's answers (real sequence: error bounds in the thousands,
    repeated 21-mers, escaped rank-line entries)."""
    import sapling_b200 as S
    g, k, nb, five, perfect, kmers, answers = load_chr3()
    for tune in ("part=0", "part_min=1,part_bits=6,chunk_log2=22"):
        monkeypatch.setenv("SAPLING_B200_TUNE", tune)
        ix = S.Sapling.from_memory(g, None, k=k)
        assert (ix.n, ix.buckets, ix.five, ix.perfectPredictions) == (len(g), nb, five, perfect)
        assert np.array_equal(ix.queryBatch(kmers), answers), tune
        ix.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")

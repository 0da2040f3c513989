"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

CPU (-m "not gpu"): the oracle port reproduces every golden fixture (this is what pins the oracle on a
box without /root/reference).  GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

import _oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith("seeds."))
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

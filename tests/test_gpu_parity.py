"""GPU parity tests: the CUDA path (through the C ABI of libsapling_b200.so) against the oracle
(oracle/sapling_oracle.c, itself pinned to the unmodified reference) on the same seeded inputs.
Bar: bit-exact (integer / index work)."""
import os
import tempfile

import numpy as np
import pytest

import _fixtures as F
import _oracle as O
from sapling_b200.api import pack_kmer_bits

pytestmark = pytest.mark.gpu

GENOMES = F.small_genomes()
CASES = [(name, k, nb) for name in GENOMES for k in (11, 16, 21, 31) for nb in (-1, 4, 10)
         if len(GENOMES[name]) >= 4 * k and nb <= 2 * k]


@pytest.fixture(scope="module")
def S():
    import sapling_b200
    return sapling_b200


@pytest.mark.parametrize("name,k,nb", CASES)
def test_query_parity_with_reference_model(S, oracle_built, name, k, nb):
    """Index parts (SA, model) from the oracle -> isolates the query kernel."""
    g = GENOMES[name]
    port = O.Port.from_memory(g, nb=nb, k=k)
    ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
    kmers = F.query_mix(g, k, 6000)
    exp, probes, oob = port.query_batch(kmers, nthreads=4, stats=True)
    got = ix.queryBatch(kmers)
    assert np.array_equal(got, exp)
    assert ix.oob_count() == oob
    # predictions alone (queryPiecewiseLinear, IEEE double, no FMA)
    pred = ix.queryPiecewiseLinear(kmers[:2000])
    assert [int(p) for p in pred] == [port.predict(int(x)) for x in kmers[:2000]]
    # variable-length strings through plQuery(s, kmerizeAdjusted, length)
    strs = F.var_len_strings(g, k, 30)
    km = [S.kmerize_adjusted(k, len(s), s) for s in strs]
    assert km == [O.kmerize_adjusted(k, len(s), s) for s in strs]
    exp_s = np.array([port.query_str(s, x) for s, x in zip(strs, km)], dtype=np.int64)
    got_s = ix.plQueryBatch(strs, km)
    assert np.array_equal(got_s, exp_s)
    # single-call surface
    for s, x, e in list(zip(strs, km, exp_s))[:5]:
        assert ix.plQuery(s, x, len(s)) == e
    ix.close()
    port.close()


@pytest.mark.parametrize("name", list(GENOMES))
def test_gpu_built_index_matches_oracle(S, oracle_built, name):
    """Suffix array, model checkpoints and error bounds built on the GPU == the oracle's; the .sa and
    .sap files are byte-identical."""
    g = GENOMES[name]
    for k, nb in ((21, -1), (16, 8), (11, -1), (31, 12)):
        if len(g) < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        ix = S.Sapling.from_memory(g, None, numBuckets=nb, k=k, flags=S.QUIET | S.KEEP_BUILD)
        assert np.array_equal(ix.rev(), port.sa), "suffix array"
        assert np.array_equal(ix.sa(), port.isa), "inverse suffix array"
        assert ix.check_sa() == (0, 0, 0)
        assert ix.buckets == port.nb
        x, y = ix.model()
        assert np.array_equal(x, port.xlist), "xlist"
        assert np.array_equal(y, port.ylist), "ylist"
        assert ix.five == port.five
        assert ix.perfectPredictions == port.perfect
        with tempfile.TemporaryDirectory() as tmp:
            a, b = os.path.join(tmp, "a.sap"), os.path.join(tmp, "b.sap")
            ix.write_sap(a)
            port.write_sap(b)
            assert open(a, "rb").read() == open(b, "rb").read(), ".sap bytes"
            a, b = os.path.join(tmp, "a.sa"), os.path.join(tmp, "b.sa")
            ix.write_sa(a)
            port.write_sa(b)
            assert open(a, "rb").read() == open(b, "rb").read(), ".sa bytes"
        # countHitsLeft/Right
        ranks = np.arange(0, len(g), max(1, len(g) // 500), dtype=np.uint32)
        left, right = ix.countHits(ranks, 32)
        exp = [port.count_hits(int(r), 32) for r in ranks]
        assert [(int(a), int(b)) for a, b in zip(left, right)] == exp
        kmers = F.query_mix(g, k, 3000, seed=5)
        assert np.array_equal(ix.queryBatch(kmers), port.query_batch(kmers, nthreads=4))
        ix.close()
        port.close()


def test_constructor_from_files(S, oracle_built, tmp_path):
    """Sapling(refFn, saFn, sapFn, ...) load-or-build semantics (sapling_api.h:492-676)."""
    text = b">chrA some description\nacgtNNacgtTTGA\nCCGATRYK\n>chrB\n" + GENOMES["rand20k"][:5000] + \
        b"\n\n>chrC x\n" + GENOMES["gc2332"][:3000] + b"\n"
    fa = tmp_path / "g.fa"
    fa.write_bytes(text)
    g_exp, ends_exp = O.clean_fasta_text(text)
    sa_fn, sap_fn = str(tmp_path / "g.fa.sa"), str(tmp_path / "g.fa.sap")
    ix = S.Sapling(str(fa), sa_fn, sap_fn, -1, -1, 16, "")
    assert ix.reference == g_exp
    assert ix.chrEnds == sorted(ends_exp)
    port = O.Port.open(str(fa), str(tmp_path / "p.sa"), str(tmp_path / "p.sap"), k=16)
    assert open(sa_fn, "rb").read() == open(tmp_path / "p.sa", "rb").read()
    assert open(sap_fn, "rb").read() == open(tmp_path / "p.sap", "rb").read()
    kmers = F.query_mix(g_exp, 16, 2000)
    exp = port.query_batch(kmers)
    assert np.array_equal(ix.queryBatch(kmers), exp)
    ix.close()
    # second open: both files exist -> loaded, not rebuilt
    ix2 = S.Sapling(str(fa), sa_fn, sap_fn, -1, -1, 16, "")
    assert ix2.buckets == port.nb and ix2.five == port.five
    assert np.array_equal(ix2.queryBatch(kmers), exp)
    ix2.close()
    port.close()


def test_errors_file(S, oracle_built, tmp_path):
    g = GENOMES["rand20k"]
    fa = tmp_path / "g.fa"
    O.write_fasta(str(fa), g)
    e1, e2 = str(tmp_path / "gpu.err"), str(tmp_path / "cpu.err")
    ix = S.Sapling(str(fa), str(tmp_path / "a.sa"), str(tmp_path / "a.sap"), 8, -1, 21, e1)
    port = O.Port.open(str(fa), str(tmp_path / "b.sa"), str(tmp_path / "b.sap"), nb=8, k=21, err_fn=e2)
    assert open(e1).read() == open(e2).read()
    ix.close()
    port.close()


def test_empty_and_tiny_batches(S, oracle_built):
    g = GENOMES["rand20k"]
    port = O.Port.from_memory(g, k=21)
    ix = S.Sapling.from_model(g, port.sa, 21, port.nb, port.xlist, port.ylist, port.five)
    assert len(ix.queryBatch(np.empty(0, dtype=np.uint64))) == 0
    one = F.query_mix(g, 21, 4)[:1]
    assert np.array_equal(ix.queryBatch(one), port.query_batch(one))
    ix.close()
    port.close()


def test_large_batch_properties_and_chunking(S, oracle_built):
    """> 1 staging chunk (2^22 queries) through the host API; every present k-mer must come back as a
    position that spells it (the self-check of sapling_example.cpp:144-154), and a 200k-query prefix
    must equal the oracle."""
    n = 3_000_000
    g = O.synth_genome(O.SEED_G + 3, n)
    ix = S.Sapling.from_memory(g, None, k=21)
    nq = (1 << 22) + 12345
    kmers, pos = O.present_queries(g, 21, nq)
    got = ix.queryBatch(kmers)
    assert (got >= 0).all()
    assert np.array_equal(O.kmers_at(g, got, 21), kmers)
    port = O.Port.from_memory(g, sa=ix.rev(), k=21)
    assert ix.five == port.five
    assert np.array_equal(got[:200000], port.query_batch(kmers[:200000], nthreads=8))
    # device-side generator and verifier agree with the host ones
    import torch
    d_k = torch.empty(100000, dtype=torch.int64, device="cuda")
    d_o = torch.empty(100000, dtype=torch.int64, device="cuda")
    ix.sample_queries_device(O.SEED_Q, 0, 0, 100000, d_k.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_k.cpu().numpy().astype(np.uint64), kmers[:100000])
    ix.queryBatchDevice(d_k.data_ptr(), 100000, d_o.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), got[:100000])
    assert ix.verify_device(d_k.data_ptr(), d_o.data_ptr(), 100000) == (100000, 0)
    ix.sample_queries_device(O.SEED_Q, O.SEED_M, 0, 100000, d_k.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_k.cpu().numpy().astype(np.uint64), O.mutate_queries(kmers[:100000], 21))
    ix.close()
    port.close()


def test_no_compat_equals_compat_below_2g(S, oracle_built):
    g = GENOMES["rand200k"]
    port = O.Port.from_memory(g, k=21)
    a = S.Sapling.from_model(g, port.sa, 21, port.nb, port.xlist, port.ylist, port.five)
    b = S.Sapling.from_model(g, port.sa, 21, port.nb, port.xlist, port.ylist, port.five, flags=S.QUIET | S.NO_COMPAT)
    kmers = F.query_mix(g, 21, 20000)
    assert np.array_equal(a.queryBatch(kmers), b.queryBatch(kmers))
    a.close(); b.close(); port.close()


def test_layout_and_hint_variants_agree(S, oracle_built, monkeypatch):
    """The narrow (8 B/bucket) and wide model layouts, every L2-hint combination and every compiled occupancy return the
    oracle's answers; genomes with many empty buckets exercise the forward-fill encoding.  (SAPLING_B200_TUNE is read
    once, when an index is created.)"""
    for name, k, nb in (("gc0110", 16, 10), ("gc1991", 21, -1), ("rand200k", 21, -1), ("tandem50", 21, 8)):
        g = GENOMES[name]
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 8000, seed=11)
        exp = port.query_batch(kmers, nthreads=4)
        for narrow, hints, occ in ((0, 0, 0), (1, 0, 3), (1, 15, 4), (0, 15, 5), (1, 5, 6), (1, 27, 0)):
            # (inorder_min=1: the pipelined in-order kernel on this small unpartitioned batch, in half the combinations)
            monkeypatch.setenv("SAPLING_B200_TUNE", f"narrow={narrow},hints={hints},occ={occ},inorder_min={1 if occ in (0, 3, 5) else -1}")
            ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
            assert np.array_equal(ix.queryBatch(kmers), exp), (name, narrow, hints, occ)
            got32 = ix.queryBatchU32(kmers)
            assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), exp), (name, "u32")
            pred = ix.queryPiecewiseLinear(kmers[:500])
            assert [int(p) for p in pred] == [port.predict(int(x)) for x in kmers[:500]]
            x, y = ix.model()  # rebuilt from the narrow table when that is the resident form
            assert np.array_equal(x, port.xlist) and np.array_equal(y, port.ylist)
            assert np.array_equal(ix.rev(), port.sa)  # read back out of the rank lines
            ix.close()
        port.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")


@pytest.mark.parametrize("name,k", [("rand200k", 32), ("rand200k", 16), ("rand200k", 21), ("gc1991", 31), ("tandem50", 16),
                                    ("repeat_tailA", 32)])
def test_answers_cross_checked_by_match_ranges(S, oracle_built, name, k):
    """Independent of any plQuery restatement: every answer of a batch in which half the k-mers carry 1-2 substitutions is
    checked against the match range [left, left + count) that the reference's libdivsufsort sa_search
    (suffixarray/libdivsufsort/lib/utils.c:259-326) returns for the same pattern -- an answer that spells the query lies
    inside the range, and a k-mer that occurs (count > 0) is found.  k = 32 has no other oracle: the reference's signed
    64-bit hash breaks there, the GPU path computes unsigned."""
    g = GENOMES[name]
    ix = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET | S.KEEP_BUILD)
    kmers, pos = O.present_queries(g, k, 20000)
    mixed = O.mutate_queries(kmers, k)          # odd entries carry 1-2 substitutions
    got = ix.queryBatch(mixed)
    sa, isa = ix.rev(), ix.sa()
    if O.divsuf_available():
        left, cnt = O.sa_search_batch(g, sa, mixed, k)
    else:  # the port's helper, pinned to sa_search where the reference tree is present (tests/test_golden.py)
        port = O.Port.from_memory(g, sa=sa, k=min(k, 21))
        rng = [port.equal_range(O.unpack_kmer(int(x), k)) for x in mixed]
        left = np.array([r[0] for r in rng], dtype=np.int64)
        cnt = np.array([r[1] - r[0] for r in rng], dtype=np.int64)
        port.close()
    inside = (got >= 0) & (got + k <= len(g))
    spelled = np.zeros(len(mixed), dtype=bool)
    spelled[inside] = O.kmers_at(g, got[inside], k) == mixed[inside]
    present = np.arange(len(mixed)) % 2 == 0
    assert spelled[present].all()                                   # sapling_example's self-check (:144-154)
    assert (spelled | (cnt == 0)).all(), "a k-mer that occurs must be found"
    rank = isa[np.clip(got, 0, len(g) - 1)].astype(np.int64)
    assert ((rank >= left) & (rank < left + cnt))[spelled].all(), "answers lie inside sa_search's match range"
    ix.close()


@pytest.mark.parametrize("name,k", [("rand200k", 16), ("rand200k", 21), ("gc1991", 16), ("tandem50", 16), ("repeat_tailA", 21),
                                    ("gc0110", 11)])
def test_seed_batch_matches_align_seed_loop(S, oracle_built, name, k):
    """sapling_b200_seed_batch == the oracle's restatement of align.cpp:259-300 (itself pinned to the reference's own
    methods, tests/test_cpu_host.py): hit positions, ranks and left/right hit counts, both strands."""
    g = GENOMES[name]
    reads, _ = O.simulate_reads(g, 300, min(150, len(g) // 4))
    reads += [b"ACGT", g[100:100 + k], g[7:7 + k + 1], b"N" * 60, g[5:155].lower(), g[-150:], g[:150]]
    port = O.Port.from_memory(g, k=k)
    ix = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET | S.KEEP_BUILD)
    for num_seeds, max_hits in ((7, 32), (1, 4), (3, 255)):
        exp = port.seed_batch(reads, num_seeds, max_hits)
        got = ix.seedBatch(reads, num_seeds, max_hits)
        for a, b, what in zip(got, exp, ("ref_pos", "sa_pos", "left", "right")):
            assert np.array_equal(a, b), (name, k, num_seeds, max_hits, what)
    assert (exp[0] >= 0).any()
    ix.close()
    port.close()


@pytest.mark.parametrize("name", ["rand200k", "gc1991", "gc0110", "tandem50", "repeat_tailA", "polyC"])
def test_rank_line_layout(S, oracle_built, name, monkeypatch):
    """The rank lines (the resident form of the suffix array) built on the GPU are byte-identical to the host build of the
    same packing code, and the k-mer path returns the oracle's answers for entry prefixes shorter than / as long as /
    longer than k (ties decided by the genome) and wide enough that 21-bit deltas overflow (escaped entries)."""
    import ctypes as C
    import test_cpu_host as T
    L = T._sim()
    g = GENOMES[name]
    n = len(g)
    for k, nb in ((21, -1), (16, 8), (11, 4), (31, 10)):
        if n < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 6000, seed=5)
        tail = np.array([O.kmerize(k, g[i:i + k] + b"A" * k) for i in range(max(0, n - 40), n)], dtype=np.uint64)
        kmers = np.concatenate([kmers, tail])
        exp, _, oob = port.query_batch(kmers, nthreads=4, stats=True)
        packed = F.pack_genome(g)
        model = np.ascontiguousarray(np.stack([port.xlist, port.ylist], axis=1).reshape(-1))
        last = np.array([port.xlist[-1], port.ylist[-1]], dtype=np.int64)
        five = np.array(port.five, dtype=np.int32)
        for bases in (0, 6, 12, 21, 31):
            monkeypatch.setenv("SAPLING_B200_TUNE", f"line_bases={bases}")
            ix = S.Sapling.from_memory(g, port.sa, numBuckets=nb, k=k)
            assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, nb, bases)
            assert ix.oob_count() == oob
            # host build of the same lines == what the simulation answers from
            out, c, _ = T._answer(L, packed, port.sa, model, n, k, port.nb, five, 1, kmers, None, last, bases)
            assert np.array_equal(out, exp)
            ix.close()
        port.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")


@pytest.mark.parametrize("name", ["rand200k", "gc0110", "tandem50", "repeat_tailA", "polyC"])
def test_partitioned_batch(S, oracle_built, name, monkeypatch):
    """The partitioned batch path (histogram, scans, staged scatter, in-order query kernel, un-permute) returns the answers
    of the unpartitioned kernel and of the oracle for every slice count, both slot formats (inside the k-mer word for
    k <= 25, side array above), ragged last chunks and 64-bit / 32-bit outputs."""
    g = GENOMES[name]
    for k, nb in ((21, -1), (16, 6), (31, 10), (11, 4)):
        if len(g) < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 40000, seed=9)[:37777]  # ragged last chunk (16384 per chunk)
        exp = port.query_batch(kmers, nthreads=4)
        for bits in (1, 3, 5, 8, 11):
            if bits > 2 * k:
                continue
            monkeypatch.setenv("SAPLING_B200_TUNE", f"part_min=1,part_bits={bits},chunk_log2=22")
            ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
            assert ix.partition_bits(len(kmers)) == bits
            assert ix.query_kernel(len(kmers))[0] == "kmer_query_ordered_kernel"
            assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, nb, bits)
            got32 = ix.queryBatchU32(kmers)
            assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), exp), (name, k, bits, "u32")
            ix.close()
        for tune, kernel in (("part=0,inorder_min=1", "kmer_query_ordered_kernel"), ("part=0", "kmer_query_kernel")):
            monkeypatch.setenv("SAPLING_B200_TUNE", tune)
            ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
            assert ix.partition_bits(len(kmers)) == 0 and ix.query_kernel(len(kmers))[0] == kernel
            assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, tune)
            ix.close()
        port.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")


def test_index_cache_round_trip(S, oracle_built, tmp_path, monkeypatch):
    """an index opened from the reference's files (two-record FASTA; .sa and .sap built and written on the
    way), saved as a private cache and restored from it is the same index: scalars, chromosome table, genome, suffix
    array, model, .sap bytes, and the answers to a mixed query batch (checked against the oracle too).  Damaged cache
    files are refused."""
    g = GENOMES["rand200k"]
    half = len(g) // 2
    fa = tmp_path / "two.fa"
    with open(fa, "wb") as f:
        for name, seq in ((b"chrA first record", g[:half]), (b"chrB", g[half:])):
            f.write(b">" + name + b"\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + b"\n")
    k = 21
    a = S.Sapling(str(fa), str(tmp_path / "two.sa"), str(tmp_path / "two.sap"), k=k, flags=S.QUIET)
    cache = tmp_path / "two.b200"
    a.save_cache(str(cache))
    b = S.Sapling.from_cache(str(cache), flags=S.QUIET)
    assert (b.n, b.k, b.buckets, b.five, b.perfectPredictions) == (a.n, a.k, a.buckets, a.five, a.perfectPredictions)
    assert b.chrEnds == a.chrEnds and len(a.chrEnds) == 2
    assert b.reference == a.reference == g
    assert np.array_equal(b.rev(), a.rev())
    xa, ya = a.model()
    xb, yb = b.model()
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    b.write_sap(str(tmp_path / "restored.sap"))
    assert open(tmp_path / "restored.sap", "rb").read() == open(tmp_path / "two.sap", "rb").read()
    kmers = F.query_mix(g, k, 20000, seed=3)
    port = O.Port.from_memory(g, k=k)
    exp = port.query_batch(kmers, nthreads=4)
    port.close()
    assert np.array_equal(a.queryBatch(kmers), exp)
    assert np.array_equal(b.queryBatch(kmers), exp)
    assert b.plQuery(g[1000:1000 + k], S.kmerize(k, g[1000:1000 + k]), k) == a.plQuery(g[1000:1000 + k], S.kmerize(k, g[1000:1000 + k]), k)
    a.close()
    b.close()
    # the reference constructor with SAPLING_B200_CACHE=1: the first open leaves <sapFn>.b200, the second one reads it
    monkeypatch.setenv("SAPLING_B200_CACHE", "1")
    c1 = S.Sapling(str(fa), str(tmp_path / "two.sa"), str(tmp_path / "two.sap"), k=k, flags=S.QUIET)
    assert os.path.exists(str(tmp_path / "two.sap") + ".b200")
    os.rename(fa, str(fa) + ".gone")  # a cached open touches none of the reference's files
    c2 = S.Sapling(str(fa), str(tmp_path / "two.sa"), str(tmp_path / "two.sap"), k=k, flags=S.QUIET)
    assert np.array_equal(c1.queryBatch(kmers), exp) and np.array_equal(c2.queryBatch(kmers), exp)
    assert c2.chrEnds == c1.chrEnds and c2.reference == g
    os.rename(str(fa) + ".gone", fa)
    c3 = S.Sapling(str(fa), str(tmp_path / "two.sa"), str(tmp_path / "k16.sap"), k=16, flags=S.QUIET)  # other k: own cache
    assert c3.k == 16 and os.path.exists(str(tmp_path / "k16.sap") + ".b200")
    for c in (c1, c2, c3):
        c.close()
    monkeypatch.delenv("SAPLING_B200_CACHE")
    raw = open(cache, "rb").read()
    for bad in (raw[:len(raw) // 2], b"X" + raw[1:], raw[:-8] + b"garbage!"):
        open(tmp_path / "bad.b200", "wb").write(bad)
        with pytest.raises(S.SaplingError):
            S.Sapling.from_cache(str(tmp_path / "bad.b200"), flags=S.QUIET)


@pytest.mark.parametrize("name,k", [("rand200k", 21), ("tandem50", 16), ("gc0110", 31)])
def test_partitioned_large_batch(S, oracle_built, name, k, monkeypatch):
    """A batch of several hundred chunks through the default partitioning rule (>= 2^22 queries) and through forced slice
    counts, packed 6-byte k-mers in and 32-bit answers out (the narrow host format), against the oracle."""
    g = GENOMES[name]
    port = O.Port.from_memory(g, k=k)
    base = F.query_mix(g, k, 300000, seed=21)
    kmers = np.tile(base, 15)[:(1 << 22) + 12345]
    exp = port.query_batch(base, nthreads=4)
    exp = np.tile(exp, 15)[:len(kmers)]
    for tune in ("", "part_bits=7", "part_bits=11,occ=4", "part=0"):
        if tune:
            monkeypatch.setenv("SAPLING_B200_TUNE", tune)
        ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
        assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, tune)
        kb = (2 * k + 7) // 8
        raw = kmers if kb == 8 else np.ascontiguousarray(kmers.view(np.uint8).reshape(-1, 8)[:, :kb]).reshape(-1)
        got32 = ix.queryBatchU32(raw, kmer_bytes=kb, nq=len(kmers))
        assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), exp), (name, k, tune, "packed")
        got32 = ix.queryBatchBits(pack_kmer_bits(kmers, 2 * k), 2 * k, len(kmers))  # nothing but the k-mers
        assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), exp), (name, k, tune, "bits")
        ix.close()
    monkeypatch.delenv("SAPLING_B200_TUNE", raising=False)
    port.close()


@pytest.mark.parametrize("name,k", [("rand200k", 21), ("gc1991", 16), ("tandem50", 31), ("rand200k", 32), ("gc0110", 11)])
def test_bit_stream_upload_format(S, oracle_built, name, k):
    """sapling_b200_query_batch_bits: k-mers as a little-endian bit stream of kmer_bits bits each -- exactly 2k, odd widths,
    whole bytes, 64 -- give the answers of the uint64 entry point, for ragged counts (streams ending inside a byte) and for
    batches of several pipeline chunks; widths below 2k are refused."""
    g = GENOMES[name]
    ix = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET)
    kmers = F.query_mix(g, k, 70001, seed=k)
    exp = ix.queryBatch(kmers)
    if k <= 31:
        port = O.Port.from_memory(g, k=k)
        assert np.array_equal(exp, port.query_batch(kmers, nthreads=4))
        port.close()
    for bits in sorted({2 * k, min(2 * k + 1, 64), min(2 * k + 7, 64), 8 * ((2 * k + 7) // 8), 64}):
        for m in (len(kmers), 1, 7, 8, 9, 4099):
            got = ix.queryBatchBits(pack_kmer_bits(kmers[:m], bits), bits, m)
            assert np.array_equal(np.where(got == 0xFFFFFFFF, -1, got.astype(np.int64)), exp[:m]), (name, k, bits, m)
    big = np.tile(kmers, 20)[: (1 << 20) + 5]
    got = ix.queryBatchBits(pack_kmer_bits(big, 2 * k), 2 * k, len(big))
    assert np.array_equal(np.where(got == 0xFFFFFFFF, -1, got.astype(np.int64)), np.tile(exp, 20)[: len(big)])
    with pytest.raises(S.SaplingError):
        ix.queryBatchBits(pack_kmer_bits(kmers[:8], 64), 2 * k - 1, 8)
    ix.close()


@pytest.mark.parametrize("name,k,nb", [("rand200k", 21, -1), ("gc1991", 16, 10), ("tandem50", 31, 8), ("repeat_tailA", 21, -1),
                                       ("polyC", 11, 4), ("gc0110", 31, -1)])
def test_inorder_kernel_on_unpartitioned_batches(S, oracle_built, name, k, nb, monkeypatch):
    """A batch that is not partitioned (index small enough for L2, or partitioning switched off) goes through the same
    pipelined in-order kernel, in the caller's order, writing the answers directly: every batch size around the tile and
    claim granularity (32 queries per tile, 128 per claim), int64 and uint32 answers, against the oracle and against the
    one-query-per-thread kernel."""
    g = GENOMES[name]
    if len(g) < 4 * k:
        pytest.skip("genome shorter than 4k")
    port = O.Port.from_memory(g, nb=nb, k=k)
    kmers = F.query_mix(g, k, 40000, seed=5)[:33333]
    exp = port.query_batch(kmers, nthreads=4)
    monkeypatch.setenv("SAPLING_B200_TUNE", "part=0,inorder_min=1")
    ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
    for m in (1, 2, 31, 32, 33, 127, 128, 129, 1000, 4097, len(kmers)):
        assert ix.query_kernel(m)[0] == "kmer_query_ordered_kernel"
        assert np.array_equal(ix.queryBatch(kmers[:m]), exp[:m]), (name, k, m)
        got32 = ix.queryBatchU32(kmers[:m])
        assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), exp[:m]), (name, k, m, "u32")
    ix.close()
    monkeypatch.setenv("SAPLING_B200_TUNE", "part=0,inorder_min=-1")
    ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
    assert ix.query_kernel(len(kmers))[0] == "kmer_query_kernel"
    assert np.array_equal(ix.queryBatch(kmers), exp)
    ix.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")
    port.close()


def test_stray_high_bits_are_ignored(S, oracle_built, monkeypatch):
    """Bits above 2k are not part of a k-mer: a word that carries them (a sign-extended or wider hash) is answered like the
    masked k-mer on every path, never used to index the model (which would read out of bounds)."""
    g = GENOMES["rand200k"]
    k = 21
    port = O.Port.from_memory(g, k=k)
    kmers = F.query_mix(g, k, 50000, seed=2)
    exp = port.query_batch(kmers, nthreads=4)
    dirty = kmers | (np.uint64(0xABCDE) << np.uint64(42)) | (np.uint64(1) << np.uint64(63))
    for tune in ("part=0", "part=0,inorder_min=1", "part_min=1,part_bits=5,chunk_log2=22"):
        monkeypatch.setenv("SAPLING_B200_TUNE", tune)
        ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
        assert np.array_equal(ix.queryBatch(dirty), exp), tune
        assert ix.oob_count() == 0
        ix.close()
    monkeypatch.delenv("SAPLING_B200_TUNE")
    port.close()


def test_strings_with_other_bytes_are_rejected(S, oracle_built):
    """A query string holding a byte other than A/C/G/T is refused: the reference compares raw bytes (an N never matches
    and sorts between G and T), which a 2-bit index cannot reproduce -- better an error than a different answer."""
    g = GENOMES["rand200k"]
    ix = S.Sapling.from_memory(g, None, k=21)
    s = bytearray(g[500:540])
    s[7] = ord("N")
    with pytest.raises(S.SaplingError):
        ix.plQuery(bytes(s), S.kmerize(21, bytes(s)), 21)
    with pytest.raises(S.SaplingError):
        ix.plQueryBatch([g[100:140], bytes(s)], [S.kmerize(21, g[100:140]), S.kmerize(21, bytes(s))])
    assert ix.plQuery(g[100:140], S.kmerize(21, g[100:140]), 21) >= 0
    ix.close()


def test_replicas_answer_like_the_primary(S, oracle_built):
    """sapling_b200_replicate: the index copied device to device onto every other GPU of the box; a host batch sharded over
    all of them returns exactly the single-GPU answers (needs >= 2 GPUs)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    g = GENOMES["rand200k"]
    k = 21
    ix = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET | S.KEEP_BUILD)
    kmers = np.tile(F.query_mix(g, k, 200000, seed=4), 8)
    one = ix.queryBatch(kmers)
    assert ix.replicate((1 << ngpu) - 1) == ngpu
    assert np.array_equal(ix.queryBatch(kmers), one)
    got32 = ix.queryBatchU32(kmers)
    assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), one)
    # bit streams are cut between the GPUs at multiples of 8 k-mers (byte boundaries); a ragged count on purpose
    m = len(kmers) - 13
    got32 = ix.queryBatchBits(pack_kmer_bits(kmers[:m], 2 * k), 2 * k, m)
    assert np.array_equal(np.where(got32 == 0xFFFFFFFF, -1, got32.astype(np.int64)), one[:m])
    ix.close()


@pytest.mark.parametrize("name,k", [("rand200k", 21), ("gc1991", 16), ("tandem50", 16), ("repeat_tailA", 21)])
def test_device_probe_count_matches_oracle(S, oracle_built, name, k):
    """SURVEY 8d's P (getLcp calls per query of the reference algorithm) counted on the device == counted by the oracle."""
    import torch
    g = GENOMES[name]
    port = O.Port.from_memory(g, k=k)
    ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
    kmers = F.query_mix(g, k, 20000)
    _, probes, _ = port.query_batch(kmers, nthreads=4, stats=True)
    d = torch.from_numpy(kmers.astype(np.int64)).cuda()
    assert ix.count_probes_device(d.data_ptr(), len(kmers)) == probes
    ix.close()
    port.close()

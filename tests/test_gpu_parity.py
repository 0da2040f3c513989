"""GPU parity tests: the CUDA path (through the C ABI of libsapling_b200.so) against the oracle
(oracle/sapling_oracle.c, itself pinned to the unmodified reference) on the same seeded inputs.
Bar: bit-exact (integer / index work)."""
import os
import tempfile

import numpy as np
import pytest

import _fixtures as F
import _oracle as O

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
    """The narrow (8 B/bucket) model layout and every L2-hint combination return the oracle's answers; genomes with
    many empty buckets exercise the forward-fill encoding."""
    for name, k, nb in (("gc0110", 16, 10), ("gc1991", 21, -1), ("rand200k", 21, -1), ("tandem50", 21, 8)):
        g = GENOMES[name]
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 8000, seed=11)
        exp = port.query_batch(kmers, nthreads=4)
        for narrow, hints in ((0, 0), (1, 0), (1, 15), (0, 15), (1, 5)):
            monkeypatch.setenv("SAPLING_B200_NARROW", str(narrow))
            monkeypatch.setenv("SAPLING_B200_HINTS", str(hints))
            ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, port.five)
            for sector, line, pipe, qv in (("1", "0", "0", "4"), ("1", "0", "0", "3"), ("1", "0", "0", "5"),
                                           ("1", "0", "0", "6"), ("0", "1", "1", "4"), ("0", "1", "1", "3"),
                                           ("0", "1", "0", "4"), ("0", "1", "0", "8"), ("0", "0", "1", "4"),
                                           ("0", "0", "1", "5"), ("0", "0", "0", "4"), ("0", "0", "0", "6"),
                                           ("0", "0", "0", "8")):
                monkeypatch.setenv("SAPLING_B200_SECTOR", sector)
                monkeypatch.setenv("SAPLING_B200_LINE", line)
                monkeypatch.setenv("SAPLING_B200_PIPELINE", pipe)
                monkeypatch.setenv("SAPLING_B200_QV", qv)
                assert np.array_equal(ix.queryBatch(kmers), exp), (name, narrow, hints, sector, line, pipe, qv)
            pred = ix.queryPiecewiseLinear(kmers[:500])
            assert [int(p) for p in pred] == [port.predict(int(x)) for x in kmers[:500]]
            ix.close()
        port.close()


def test_k32_cross_checked_by_equal_range(S, oracle_built):
    """k=32 has no reference oracle (the reference's signed 64-bit hash breaks there): unsigned arithmetic on
    the GPU, validated against the independent match-range oracle (the contract of libdivsufsort's sa_search)."""
    g = GENOMES["rand200k"]
    k = 32
    ix = S.Sapling.from_memory(g, None, k=k)
    kmers, pos = O.present_queries(g, k, 20000)
    mixed = O.mutate_queries(kmers, k)          # odd entries carry 1-2 substitutions
    got = ix.queryBatch(mixed)
    port = O.Port.from_memory(g, sa=ix.rev(), k=21)  # only its suffix array / equal_range are used
    isa = port.isa
    spelled = O.kmers_at(g, np.clip(got, 0, len(g) - k), k)
    present = np.arange(len(mixed)) % 2 == 0
    assert (got[present] >= 0).all() and np.array_equal(spelled[present], mixed[present])   # sapling_example's self-check
    for i in list(range(0, 2000, 2)) + list(range(1, 2000, 2)):
        lb, ub = port.equal_range(O.unpack_kmer(int(mixed[i]), k))
        if got[i] >= 0 and got[i] + k <= len(g) and spelled[i] == mixed[i]:
            assert lb <= isa[got[i]] < ub
        else:
            assert i % 2 == 1 and lb == ub, "a k-mer that occurs must be found"
    assert ix.verify_device is not None
    ix.close()
    port.close()


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
    for num_seeds, max_hits in ((7, 32), (1, 4), (3, 1000)):
        exp = port.seed_batch(reads, num_seeds, max_hits)
        got = ix.seedBatch(reads, num_seeds, max_hits)
        for a, b, what in zip(got, exp, ("ref_pos", "sa_pos", "left", "right")):
            assert np.array_equal(a, b), (name, k, num_seeds, max_hits, what)
    assert (exp[0] >= 0).any()
    ix.close()
    port.close()


@pytest.mark.parametrize("name", ["rand200k", "gc1991", "tandem50", "repeat_tailA", "polyC"])
def test_inline_prefix_suffix_array(S, oracle_built, name):
    """The inline-prefix suffix array (rank -> {position, leading bases}; default for genomes >= 400 Mbp) returns the
    oracle's answers on both build paths: from the GPU suffix-array builder's sort keys (27 bases) and by gather for a
    suffix array that was supplied (32 bases)."""
    g = GENOMES[name]
    for k, nb in ((21, -1), (16, 8), (11, -1), (27, 10), (31, 12)):
        if len(g) < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 6000, seed=3)
        exp = port.query_batch(kmers, nthreads=4)
        a = S.Sapling.from_memory(g, None, numBuckets=nb, k=k, flags=S.QUIET | S.INLINE)        # sort-key prefixes
        b = S.Sapling.from_memory(g, port.sa, numBuckets=nb, k=k, flags=S.QUIET | S.INLINE)     # gathered prefixes
        c = S.Sapling.from_memory(g, None, numBuckets=nb, k=k, flags=S.QUIET | S.NO_INLINE)
        assert a.device_bytes() >= c.device_bytes() + 16 * len(g)
        for ix in (a, b, c):
            assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, nb)
            ix.close()
        port.close()


@pytest.mark.parametrize("name", ["rand200k", "gc1991", "gc0110", "tandem50", "repeat_tailA", "polyC"])
def test_rank_line_layout(S, oracle_built, name, monkeypatch):
    """The rank-line layout (one 32-byte sector = four {position, prefix} entries; default for genomes >= 50 Mbp)
    returns the oracle's answers: overlapping (shift 3) and tiling (shift 4) lines, prefixes shorter than / equal to /
    longer than k (the first falls back to the packed genome on ties, the last overflows the 21-bit deltas and
    escapes), every resident-blocks variant of the kernel, and suffixes at the very end of the text."""
    g = GENOMES[name]
    for k, nb in ((21, -1), (16, 8), (11, -1), (31, 12)):
        if len(g) < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        kmers = F.query_mix(g, k, 6000, seed=7)
        tail = np.array([O.kmerize(k, g[i:i + k] + b"A" * k) for i in range(len(g) - 40, len(g))], dtype=np.uint64)
        kmers = np.concatenate([kmers, tail])
        exp = port.query_batch(kmers, nthreads=4)
        plain = S.Sapling.from_memory(g, port.sa, numBuckets=nb, k=k, flags=S.QUIET | S.NO_PACKED | S.NO_INLINE)
        assert plain.query_kernel()[0] == "kmer_query_sector_kernel"
        for shift in ("3", "4"):
            for bases in ("0", "8", "14", "32"):   # 0 = the default for this genome size
                monkeypatch.setenv("SAPLING_B200_PACKED_SHIFT", shift)
                monkeypatch.setenv("SAPLING_B200_PACKED_BASES", bases)
                ix = S.Sapling.from_memory(g, port.sa if bases != "8" else None, numBuckets=nb, k=k,
                                           flags=S.QUIET | S.PACKED)
                assert ix.device_bytes() >= plain.device_bytes() + (16 if shift == "3" else 8) * len(g)
                for refill, kernel in (("1", "kmer_query_packed_refill_kernel"), ("0", "kmer_query_packed_kernel")):
                    monkeypatch.setenv("SAPLING_B200_REFILL", refill)
                    # the refill kernel reads the narrow model layout, which needs 2k - nb <= 31
                    assert ix.query_kernel()[0] == (kernel if 2 * k - port.nb <= 31 else "kmer_query_packed_kernel")
                    for qv in ("2", "3", "4", "5", "6"):
                        monkeypatch.setenv("SAPLING_B200_QV", qv)
                        assert np.array_equal(ix.queryBatch(kmers), exp), (name, k, nb, shift, bases, refill, qv)
                        # ragged batch sizes: fewer queries than lanes, than warps, one over a block boundary
                        for m in (1, 31, 33, 257, 4097):
                            assert np.array_equal(ix.queryBatch(kmers[:m]), exp[:m]), (name, k, m, refill, qv)
                    monkeypatch.delenv("SAPLING_B200_QV")
                monkeypatch.delenv("SAPLING_B200_REFILL")
                ix.close()
        plain.close()
        port.close()


@pytest.mark.parametrize("name", ["rand200k", "gc0110", "tandem50", "repeat_tailA", "polyC"])
def test_partitioned_batch(S, oracle_built, name, monkeypatch):
    """The partitioned batch path (partition.cu: bucket the batch by the top bits of the k-mer, answer it slice by
    slice, put the answers back in the caller's order) returns exactly what the oracle returns for the caller's order:
    every layout, bin counts from 2 to 2048 (more bins than buckets included), batches smaller than one partition
    chunk, ragged last chunks, several chunks, k-mers that all fall into one bin."""
    g = GENOMES[name]
    for k, nb in ((21, -1), (16, 8), (11, 4), (31, 12)):
        if len(g) < 4 * k:
            continue
        port = O.Port.from_memory(g, nb=nb, k=k)
        base = F.query_mix(g, k, 6000, seed=11)
        kmers = np.concatenate([base, base[::-1], np.sort(base), np.full(3000, base[0], dtype=np.uint64)] * 2)  # 42000 > 2 chunks
        exp = port.query_batch(kmers, nthreads=4)
        for flags in (S.NO_PACKED | S.NO_INLINE, S.PACKED, S.INLINE | S.NO_PACKED):
            ix = S.Sapling.from_memory(g, port.sa, numBuckets=nb, k=k, flags=S.QUIET | flags)
            monkeypatch.setenv("SAPLING_B200_PART", "0")
            assert ix.partition_bits(len(kmers)) == 0
            plain = ix.queryBatch(kmers)
            assert np.array_equal(plain, exp)
            monkeypatch.setenv("SAPLING_B200_PART", "1")
            monkeypatch.setenv("SAPLING_B200_PART_MIN", "1")
            for bits in (1, 3, 8, 9, 10, 11):  # un-permute: run-per-warp up to 8, then lane groups of 16, 8, 4
                monkeypatch.setenv("SAPLING_B200_PART_BITS", str(bits))
                assert ix.partition_bits(len(kmers)) == min(bits, 2 * k)
                for m in (len(kmers), 8192, 8193, 16384, 16385, 32767, 1, 33, 5000):
                    assert np.array_equal(ix.queryBatch(kmers[:m]), exp[:m]), (name, k, nb, flags, bits, m)
            monkeypatch.delenv("SAPLING_B200_PART_BITS")
            assert ix.oob_count() >= 0
            ix.close()
        port.close()
    monkeypatch.delenv("SAPLING_B200_PART_MIN", raising=False)
    monkeypatch.delenv("SAPLING_B200_PART", raising=False)


@pytest.mark.gpu
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


@pytest.mark.gpu
@pytest.mark.parametrize("name,k", [("rand200k", 21), ("tandem50", 16), ("gc0110", 31)])
def test_partitioned_large_batch(S, oracle_built, name, k, monkeypatch):
    """A batch of many partition chunks (2.1 M queries = 129 chunks: the column scan of the offset table works in 64 row
    segments of more than one row, the last chunk is ragged) answered through the partitioned path, with the slot inside
    the k-mer word (k <= 25) and in the side array (k = 31), equals the same batch answered in the caller's order, and
    its first 40 000 answers equal the oracle's."""
    g = GENOMES[name]
    port = O.Port.from_memory(g, k=k)
    rng = np.random.default_rng(17)
    q0 = F.query_mix(g, k, 40000, seed=23)
    kmers = np.concatenate([q0, rng.choice(q0, size=2_100_000 - len(q0) + 777)])
    exp_head = port.query_batch(q0, nthreads=4)
    for flags in (S.PACKED, S.NO_PACKED | S.NO_INLINE):
        ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, list(port.five), flags=S.QUIET | flags)
        monkeypatch.setenv("SAPLING_B200_PART", "0")
        plain = ix.queryBatch(kmers)
        assert np.array_equal(plain[:len(q0)], exp_head)
        monkeypatch.setenv("SAPLING_B200_PART", "1")
        monkeypatch.setenv("SAPLING_B200_PART_MIN", "1")
        monkeypatch.setenv("SAPLING_B200_CHUNK_LOG2", "22")  # the host entry point hands the whole batch to one call
        for bits in (5, 10):
            monkeypatch.setenv("SAPLING_B200_PART_BITS", str(bits))
            assert ix.partition_bits(len(kmers)) == bits
            got = ix.queryBatch(kmers)
            assert np.array_equal(got, plain), (name, k, flags, bits, int((got != plain).sum()))
        ix.close()
    port.close()
    for v in ("PART", "PART_MIN", "PART_BITS", "CHUNK_LOG2"):
        monkeypatch.delenv("SAPLING_B200_" + v, raising=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rand200k", "gc0110", "tandem50"])
def test_replay_and_partition_variants_agree(S, oracle_built, name, monkeypatch):
    """Every selectable variant of the batch path returns the oracle's answers: the lean 32-bit replay against the
    general one, the anchor line staged in shared memory against sector-by-sector fetches, the in-order tile schedule
    with and without the software pipeline against the static grid-stride schedule, the staged against the direct
    scatter, the flat and the lane-group against the run-per-warp un-permute, the flat replay against the lean one -- on all three index layouts, collapsed left windows
     included."""
    g = GENOMES[name]
    for k, nb, five_fn in ((21, -1, None), (16, 8, lambda f: [f[0], f[1], f[2], f[3], 1 << 30]), (31, 12, None)):
        if len(g) < 4 * k:
            continue
        base = O.Port.from_memory(g, nb=nb, k=k)
        five = list(base.five) if five_fn is None else five_fn(list(base.five))
        port = O.Port.from_parts(g, base.sa, k, base.nb, base.xlist, base.ylist, five)
        q0 = F.query_mix(g, k, 6000, seed=5)
        kmers = np.concatenate([q0, np.sort(q0), q0[::-1]] * 2)  # 36000: several partition chunks
        exp = port.query_batch(kmers, nthreads=4)
        for flags in (S.NO_PACKED | S.NO_INLINE, S.PACKED, S.INLINE | S.NO_PACKED):
            for shift in (("3", "4") if flags == S.PACKED else ("3",)):
                monkeypatch.setenv("SAPLING_B200_PACKED_SHIFT", shift)
                ix = S.Sapling.from_model(g, port.sa, k, port.nb, port.xlist, port.ylist, five, flags=S.QUIET | flags)
                for part in ("0", "1"):
                    monkeypatch.setenv("SAPLING_B200_PART", part)
                    monkeypatch.setenv("SAPLING_B200_PART_MIN", "1")
                    monkeypatch.setenv("SAPLING_B200_PART_BITS", "5")
                    for lean, line_smem in (("1", "1"), ("1", "0"), ("0", "0")):
                        monkeypatch.setenv("SAPLING_B200_LEAN", lean)
                        monkeypatch.setenv("SAPLING_B200_LINE_SMEM", line_smem)
                        # last two: SLOT_IN_KMER (slot inside the partitioned k-mer word, k <= 25), PARK (unfinished
                        # queries parked after three probes; needs the former)
                        combos = ((("1", "1", "1", "1", "1", "1", "1"), ("1", "0", "1", "2", "1", "1", "1"),
                                   ("0", "1", "0", "0", "1", "1", "1"), ("0", "1", "1", "0", "0", "1", "1"),
                                   ("1", "1", "1", "2", "0", "1", "0"), ("1", "1", "1", "2", "0", "0", "1"),
                                   ("1", "1", "1", "1", "0", "1", "1"))
                                  if part == "1" else (("1", "1", "1", "1", "1", "1", "1"), ("1", "1", "1", "1", "0", "1", "1")))
                        for tiles, pipe, scat, unp, flat, sik, park in combos:
                            monkeypatch.setenv("SAPLING_B200_SLOT_IN_KMER", sik)
                            monkeypatch.setenv("SAPLING_B200_PARK", park)
                            monkeypatch.setenv("SAPLING_B200_FLAT", flat)  # kmer_replay_flat / kmer_replay32 (tiling lines)
                            monkeypatch.setenv("SAPLING_B200_PART_TILES", tiles)
                            monkeypatch.setenv("SAPLING_B200_ORDERED_PIPE", pipe)
                            monkeypatch.setenv("SAPLING_B200_PART_SCATTER", scat)
                            monkeypatch.setenv("SAPLING_B200_PART_UNPERMUTE", unp)
                            for qv in ("3", "4", "5"):
                                monkeypatch.setenv("SAPLING_B200_QV", qv)
                                got = ix.queryBatch(kmers)
                                assert np.array_equal(got, exp), (name, k, flags, shift, part, lean, line_smem, tiles,
                                                                  pipe, scat, unp, flat, sik, park, qv)
                            assert np.array_equal(ix.queryBatch(kmers[:8191]), exp[:8191])
                ix.close()
        port.close()
        base.close()
    for v in ("PART", "PART_MIN", "PART_BITS", "LEAN", "LINE_SMEM", "PART_TILES", "ORDERED_PIPE", "PART_SCATTER",
              "PART_UNPERMUTE", "QV", "PACKED_SHIFT", "FLAT", "SLOT_IN_KMER", "PARK"):
        monkeypatch.delenv("SAPLING_B200_" + v, raising=False)


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

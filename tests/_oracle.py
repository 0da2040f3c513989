"""ctypes bindings to the CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Port``  -> oracle/_build/liboracle.so      (oracle/sapling_oracle.c, our restatement)
* ``Ref``   -> oracle/_ref/libsapling_ref.so   (the unmodified reference header, oracle/ref_harness.cpp)

Nothing under sapling_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsapling_ref.so")
REF_UNDEFINED = -(1 << 63)  # oracle/ref_harness.cpp: the reference's answer is undefined (predicted rank >= n)

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(port=True, ref=True):
    """(Re)build the checkers with oracle/Makefile.  `ref` is a no-op without /root/reference."""
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")] + targets, check=True)


class _SoIndex(C.Structure):
    _fields_ = [
        ("ref", C.c_void_p), ("n", C.c_uint64),
        ("k", C.c_int), ("nb", C.c_int), ("maxMem", C.c_int),
        ("rev", C.POINTER(C.c_uint32)), ("inv", C.POINTER(C.c_uint32)), ("lcp", C.POINTER(C.c_uint32)),
        ("krmqb", C.POINTER(C.c_uint32)),
        ("xlist", C.POINTER(C.c_int64)), ("ylist", C.POINTER(C.c_int64)),
        ("maxOver", C.c_int), ("maxUnder", C.c_int), ("meanError", C.c_int),
        ("mostOver", C.c_int), ("mostUnder", C.c_int),
        ("perfect", C.c_uint64), ("nOver", C.c_uint64), ("nUnder", C.c_uint64),
        ("chrEndPos", C.POINTER(C.c_uint64)), ("chrEndName", C.POINTER(C.c_char_p)), ("nChr", C.c_size_t),
    ]


_port = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build(port=True, ref=False)
        L = C.CDLL(PORT_SO)
        P = C.POINTER(_SoIndex)
        L.so_kmerize.restype = C.c_int64
        L.so_kmerize.argtypes = [C.c_int, C.c_char_p]
        L.so_kmerize_adjusted.restype = C.c_int64
        L.so_kmerize_adjusted.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.so_open.restype = P
        L.so_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.so_from_memory.restype = P
        L.so_from_memory.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.so_from_parts.restype = P
        L.so_from_parts.argtypes = [C.c_char_p, C.c_uint64, u32p, C.c_int, C.c_int, i64p, i64p,
                                    np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
        L.so_close.argtypes = [P]
        L.so_predict.restype = C.c_uint64
        L.so_predict.argtypes = [P, C.c_int64]
        L.so_plquery.restype = C.c_int64
        L.so_plquery.argtypes = [P, C.c_char_p, C.c_size_t, C.c_int64, C.c_size_t,
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.so_query_batch.argtypes = [P, u64p, C.c_size_t, i64p, C.c_int,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.so_query_batch_timed.restype = C.c_double
        L.so_query_batch_timed.argtypes = [P, u64p, C.c_size_t, i64p, C.c_int]
        L.so_count_hits_left.restype = C.c_uint64
        L.so_count_hits_left.argtypes = [P, C.c_uint64, C.c_uint64]
        L.so_count_hits_right.restype = C.c_uint64
        L.so_count_hits_right.argtypes = [P, C.c_uint64, C.c_uint64]
        L.so_seed_batch.argtypes = [P, C.c_char_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, i64p, u32p, u32p, u32p, C.c_int]
        L.so_equal_range.argtypes = [P, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.so_read_sap_file.argtypes = [P, C.c_char_p]
        L.so_write_sap_file.argtypes = [P, C.c_char_p]
        L.so_write_sa_file.argtypes = [P, C.c_char_p]
        L.so_build_sap.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.so_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]
        L.so_splitmix64.restype = C.c_uint64
        L.so_splitmix64.argtypes = [C.c_uint64]
        L.so_clean_fasta_text.restype = C.c_void_p
        L.so_clean_fasta_text.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), P]
        _port = L
    return _port


def _b(s):
    return s.encode() if isinstance(s, str) else s


class Port:
    """oracle/sapling_oracle.c index."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle port: open failed")
        self.h = handle
        self.L = port_lib()

    @classmethod
    def open(cls, fa, sa, sap, nb=-1, maxMem=-1, k=-1, err_fn=None):
        return cls(port_lib().so_open(_b(fa), _b(sa), _b(sap), nb, maxMem, k, _b(err_fn) if err_fn else None))

    @classmethod
    def from_memory(cls, genome: bytes, sa=None, nb=-1, maxMem=-1, k=-1):
        sap = None
        if sa is not None:
            sa = np.ascontiguousarray(sa, dtype=np.uint32)
            sap = sa.ctypes.data_as(C.c_void_p)
        return cls(port_lib().so_from_memory(genome, len(genome), sap, nb, maxMem, k))

    @classmethod
    def from_parts(cls, genome: bytes, sa, k, nb, xlist, ylist, five):
        return cls(port_lib().so_from_parts(genome, len(genome), np.ascontiguousarray(sa, dtype=np.uint32), k, nb,
                                            np.ascontiguousarray(xlist, dtype=np.int64),
                                            np.ascontiguousarray(ylist, dtype=np.int64),
                                            np.ascontiguousarray(five, dtype=np.int32)))

    def close(self):
        if self.h:
            self.L.so_close(self.h)
            self.h = None

    # --- fields ---
    @property
    def n(self): return int(self.h.contents.n)
    @property
    def k(self): return int(self.h.contents.k)
    @property
    def nb(self): return int(self.h.contents.nb)
    @property
    def five(self):
        c = self.h.contents
        return (c.maxOver, c.maxUnder, c.meanError, c.mostOver, c.mostUnder)
    @property
    def perfect(self): return int(self.h.contents.perfect)
    @property
    def genome(self): return C.string_at(self.h.contents.ref, self.n)
    def _arr(self, ptr, count, dt):
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt, copy=True)
    @property
    def sa(self): return self._arr(self.h.contents.rev, self.n, np.uint32)
    @property
    def isa(self): return self._arr(self.h.contents.inv, self.n, np.uint32)
    @property
    def lcp(self): return self._arr(self.h.contents.lcp, self.n - 1, np.uint32)
    @property
    def xlist(self): return self._arr(self.h.contents.xlist, (1 << self.nb) + 1, np.int64)
    @property
    def ylist(self): return self._arr(self.h.contents.ylist, (1 << self.nb) + 1, np.int64)
    @property
    def chr_ends(self):
        c = self.h.contents
        return [(int(c.chrEndPos[i]), c.chrEndName[i].decode()) for i in range(c.nChr)]

    # --- queries ---
    def predict(self, x): return int(self.L.so_predict(self.h, int(x)))

    def query_str(self, s, kmer, length=None, slen=None, want_probes=False):
        s = _b(s)
        slen = len(s) if slen is None else slen
        length = slen if length is None else length
        p, f = C.c_uint32(0), C.c_uint32(0)
        r = int(self.L.so_plquery(self.h, s, slen, int(kmer), length, C.byref(p), C.byref(f)))
        return (r, p.value, f.value) if want_probes else r

    def query_batch(self, kmers, nthreads=1, stats=False):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(len(kmers), dtype=np.int64)
        pt, oob = C.c_uint64(0), C.c_uint64(0)
        self.L.so_query_batch(self.h, kmers, len(kmers), out, nthreads, C.byref(pt), C.byref(oob))
        return (out, pt.value, oob.value) if stats else out

    def query_batch_timed(self, kmers, nthreads=1):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(len(kmers), dtype=np.int64)
        t = self.L.so_query_batch_timed(self.h, kmers, len(kmers), out, nthreads)
        return out, float(t)

    def equal_range(self, s):
        s = _b(s)
        lb, ub = C.c_uint64(0), C.c_uint64(0)
        self.L.so_equal_range(self.h, s, len(s), C.byref(lb), C.byref(ub))
        return lb.value, ub.value

    def count_hits(self, sa_pos, max_hits):
        return (int(self.L.so_count_hits_left(self.h, sa_pos, max_hits)),
                int(self.L.so_count_hits_right(self.h, sa_pos, max_hits)))

    def seed_batch(self, reads, num_seeds=7, max_hits=32, nthreads=4):
        """align.cpp:259-300 seed lookups: (ref_pos, sa_pos, left, right), each of shape (n_reads, 2, num_seeds)."""
        blob, off = pack_reads(reads)
        m = len(reads) * 2 * num_seeds
        rp, sp, lf, rt = (np.empty(m, np.int64), np.empty(m, np.uint32), np.empty(m, np.uint32), np.empty(m, np.uint32))
        self.L.so_seed_batch(self.h, blob, off, len(reads), num_seeds, max_hits, rp, sp, lf, rt, nthreads)
        shp = (len(reads), 2, num_seeds)
        return rp.reshape(shp), sp.reshape(shp), lf.reshape(shp), rt.reshape(shp)

    def write_sap(self, path): return self.L.so_write_sap_file(self.h, _b(path))
    def write_sa(self, path): return self.L.so_write_sa_file(self.h, _b(path))


def pack_reads(reads):
    """Concatenated ASCII + uint64 offsets (n_reads + 1)."""
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return b"".join(reads), off


def kmerize(k, s): return int(port_lib().so_kmerize(k, _b(s)))
def kmerize_adjusted(k, length, s): return int(port_lib().so_kmerize_adjusted(k, length, _b(s)))


def synth_genome(seed, n) -> bytes:
    buf = C.create_string_buffer(n)
    port_lib().so_synth_genome(seed, n, buf)
    return buf.raw[:n]


def splitmix64(z):
    return int(port_lib().so_splitmix64(z & 0xFFFFFFFFFFFFFFFF))


def clean_fasta_text(text: bytes):
    L = port_lib()
    ix = _SoIndex()
    n = C.c_uint64(0)
    p = L.so_clean_fasta_text(text, len(text), C.byref(n), C.byref(ix))
    g = C.string_at(p, n.value)
    ends = [(int(ix.chrEndPos[i]), ix.chrEndName[i].decode()) for i in range(ix.nChr)]
    return g, ends


# ----------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_from_parts.restype = C.c_void_p
        L.ref_from_parts.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ref_parts_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_char_p]
        L.ref_parts_rev.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, u32p, C.c_int]
        L.ref_parts_model.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, i64p, i64p]
        L.ref_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        for nm in ("ref_genome", "ref_xlist", "ref_ylist", "ref_rev", "ref_inv", "ref_lcp"):
            getattr(L, nm).restype = C.c_void_p
            getattr(L, nm).argtypes = [C.c_void_p]
        L.ref_num_chr.restype = C.c_size_t
        L.ref_num_chr.argtypes = [C.c_void_p]
        L.ref_chr.restype = C.c_size_t
        L.ref_chr.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.ref_kmerize.restype = C.c_longlong
        L.ref_kmerize.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_kmerize_adjusted.restype = C.c_longlong
        L.ref_kmerize_adjusted.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.ref_predict.restype = C.c_size_t
        L.ref_predict.argtypes = [C.c_void_p, C.c_longlong]
        L.ref_query_str.restype = C.c_longlong
        L.ref_query_str.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_longlong, C.c_size_t]
        L.ref_query_batch.restype = C.c_double
        L.ref_query_batch.argtypes = [C.c_void_p, u64p, C.c_size_t, i64p, C.c_int]
        L.ref_count_hits_left.restype = C.c_size_t
        L.ref_count_hits_left.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_count_hits_right.restype = C.c_size_t
        L.ref_count_hits_right.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_seed_batch.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, i64p, u32p, u32p, u32p,
                                     C.c_int]
        L.ref_max_threads.restype = C.c_int
        _ref = L
    return _ref


class Ref:
    """The unmodified reference `struct Sapling` (sapling_api.h:17) behind oracle/ref_harness.cpp."""

    def __init__(self, fa, sa, sap, nb=-1, maxMem=-1, k=-1, err_fn=None, quiet=True, _handle=None):
        self.L = ref_lib()
        self.h = _handle if _handle is not None else self.L.ref_open(
            _b(fa), _b(sa), _b(sap), nb, maxMem, k, _b(err_fn) if err_fn else None, 1 if quiet else 0)
        n, kk, nbb, perfect = C.c_uint64(0), C.c_int(0), C.c_int(0), C.c_uint64(0)
        five = (C.c_int * 5)()
        self.L.ref_info(self.h, C.byref(n), C.byref(kk), C.byref(nbb), five, C.byref(perfect))
        self.n, self.k, self.nb = n.value, kk.value, nbb.value
        self.five = tuple(five)
        self.perfect = perfect.value

    @classmethod
    def from_parts(cls, genome: bytes, rev, k, nb, xlist, ylist, five, nthreads=4, chunk=1 << 26):
        """`struct Sapling` filled member by member (oracle/ref_harness.cpp ref_from_parts): no constructor, no files.
        `rev` is a uint32 array (rank -> position) or a callable rev(first, count) -> uint32 array, so that a 3.1 Gbp
        suffix array can be streamed from the GPU index without a second 12 GB host copy."""
        L = ref_lib()
        n = len(genome)
        f5 = (C.c_int * 5)(*[int(v) for v in five])
        h = L.ref_from_parts(n, int(k), int(nb), f5)
        if not h:
            raise RuntimeError("ref_from_parts failed")
        L.ref_parts_genome(h, 0, n, genome)
        get = rev if callable(rev) else (lambda first, count: rev[first:first + count])
        for first in range(0, n, chunk):
            c = min(chunk, n - first)
            L.ref_parts_rev(h, first, c, np.ascontiguousarray(get(first, c), dtype=np.uint32), nthreads)
        xs = np.ascontiguousarray(xlist, dtype=np.int64)
        ys = np.ascontiguousarray(ylist, dtype=np.int64)
        assert len(xs) == (1 << nb) + 1 and len(ys) == len(xs)
        L.ref_parts_model(h, 0, len(xs), xs, ys)
        return cls(None, None, None, _handle=h)

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def _arr(self, fn, count, dt):
        p = getattr(self.L, fn)(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64 if dt == np.uint64 else C.c_int64)),
                                     shape=(count,)).astype(dt, copy=True)

    @property
    def genome(self): return C.string_at(self.L.ref_genome(self.h), self.n)
    @property
    def xlist(self): return self._arr("ref_xlist", (1 << self.nb) + 1, np.int64)
    @property
    def ylist(self): return self._arr("ref_ylist", (1 << self.nb) + 1, np.int64)
    @property
    def sa(self): return self._arr("ref_rev", self.n, np.uint64)
    @property
    def isa(self): return self._arr("ref_inv", self.n, np.uint64)
    @property
    def lcp(self): return self._arr("ref_lcp", self.n - 1, np.uint64)
    @property
    def chr_ends(self):
        out = []
        buf = C.create_string_buffer(256)
        for i in range(self.L.ref_num_chr(self.h)):
            pos = self.L.ref_chr(self.h, i, buf, 256)
            out.append((int(pos), buf.value.decode()))
        return out

    def kmerize(self, s): return int(self.L.ref_kmerize(self.h, _b(s)))
    def kmerize_adjusted(self, length, s): return int(self.L.ref_kmerize_adjusted(self.h, length, _b(s)))
    def predict(self, x): return int(self.L.ref_predict(self.h, int(x)))

    def query_str(self, s, kmer, length=None, slen=None):
        s = _b(s)
        slen = len(s) if slen is None else slen
        length = slen if length is None else length
        return int(self.L.ref_query_str(self.h, s, slen, int(kmer), length))

    def query_batch(self, kmers, nthreads=1, timed=False):
        """plQuery per k-mer.  A k-mer whose predicted rank is >= n (undefined in the reference: it reads rev[] out of
        bounds) is not run; its answer is REF_UNDEFINED."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(len(kmers), dtype=np.int64)
        t = self.L.ref_query_batch(self.h, kmers, len(kmers), out, nthreads)
        return (out, float(t)) if timed else out

    def seed_batch(self, reads, num_seeds=7, max_hits=32, nthreads=4):
        blob, off = pack_reads(reads)
        m = len(reads) * 2 * num_seeds
        rp, sp, lf, rt = (np.empty(m, np.int64), np.empty(m, np.uint32), np.empty(m, np.uint32), np.empty(m, np.uint32))
        self.L.ref_seed_batch(self.h, blob, off, len(reads), num_seeds, max_hits, rp, sp, lf, rt, nthreads)
        shp = (len(reads), 2, num_seeds)
        return rp.reshape(shp), sp.reshape(shp), lf.reshape(shp), rt.reshape(shp)

    def count_hits(self, sa_pos, max_hits):
        return (int(self.L.ref_count_hits_left(self.h, sa_pos, max_hits)),
                int(self.L.ref_count_hits_right(self.h, sa_pos, max_hits)))


# ----------------------------------------------------------------------------------------------
# The reference's libdivsufsort submodule (oracle/_ref/libdivsuf_ref.so, oracle/ref_divsuf_harness.c): divsufsort() and the
# independent match-range oracle sa_search() (suffixarray/libdivsufsort/lib/utils.c:259-326).
DIVSUF_SO = os.path.join(ROOT, "oracle", "_ref", "libdivsuf_ref.so")
_divsuf = None
i64c = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def divsuf_available():
    return os.path.exists(DIVSUF_SO)


def divsuf_lib():
    global _divsuf
    if _divsuf is None:
        L = C.CDLL(DIVSUF_SO)
        L.ref_sa_search_batch.argtypes = [u8p, C.c_int64, u8p, C.c_int64, C.c_int64, i64c, i64c, i64c]
        L.ref_divsufsort.restype = C.c_int
        L.ref_divsufsort.argtypes = [u8p, i64c, C.c_int64]
        L.ref_sufcheck.restype = C.c_int
        L.ref_sufcheck.argtypes = [u8p, i64c, C.c_int64]
        _divsuf = L
    return _divsuf


def sa_search_batch(genome: bytes, sa, kmers, k):
    """(left, count) of libdivsufsort's sa_search for each packed k-mer: the ranks [left, left + count) are exactly the
    suffixes that start with it."""
    T = np.frombuffer(genome, dtype=np.uint8)
    SA = np.ascontiguousarray(sa, dtype=np.int64)
    kmers = np.asarray(kmers, dtype=np.uint64)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    pat = np.empty((len(kmers), k), dtype=np.uint8)
    for j in range(k):
        pat[:, j] = letters[((kmers >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.int64)]
    left = np.empty(len(kmers), dtype=np.int64)
    cnt = np.empty(len(kmers), dtype=np.int64)
    divsuf_lib().ref_sa_search_batch(np.ascontiguousarray(T), len(T), np.ascontiguousarray(pat).reshape(-1), k, len(kmers), SA,
                                     left, cnt)
    return left, cnt


def divsufsort(genome: bytes):
    T = np.ascontiguousarray(np.frombuffer(genome, dtype=np.uint8))
    SA = np.empty(len(T), dtype=np.int64)
    rc = divsuf_lib().ref_divsufsort(T, SA, len(T))
    assert rc == 0
    return SA


def sufcheck(genome: bytes, sa):
    T = np.ascontiguousarray(np.frombuffer(genome, dtype=np.uint8))
    return int(divsuf_lib().ref_sufcheck(T, np.ascontiguousarray(sa, dtype=np.int64), len(T)))


# ----------------------------------------------------------------------------------------------
# Shared input generators (seeded, identical everywhere)

SEED_G = 0x5A91_1C0D_E5EE_D001
SEED_Q = 0x5A91_1C0D_E5EE_D002
SEED_M = 0x5A91_1C0D_E5EE_D003

_CODE = np.zeros(256, dtype=np.uint64)
_CODE[ord("C")] = 1
_CODE[ord("G")] = 2
_CODE[ord("T")] = 3


def splitmix64_np(z):
    z = (np.asarray(z, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def kmers_at(genome: bytes, pos, k):
    """Packed k-mers (kmerize values) of genome[pos:pos+k] for an array of positions."""
    g = np.frombuffer(genome, dtype=np.uint8)  # gathered first, coded second: no 8-bytes-per-base temporary
    pos = np.asarray(pos, dtype=np.int64)
    x = np.zeros(len(pos), dtype=np.uint64)
    for j in range(k):
        x = (x << np.uint64(2)) | _CODE[g[pos + j]]
    return x


def present_queries(genome: bytes, k, nq, seed=SEED_Q):
    """pos_j = splitmix64(seed + j) mod (n-k)   """
    n = len(genome)
    with np.errstate(over="ignore"):
        pos = splitmix64_np(np.uint64(seed) + np.arange(nq, dtype=np.uint64)) % np.uint64(n - k)
    return kmers_at(genome, pos.astype(np.int64), k), pos.astype(np.int64)


def mutate_queries(kmers, k, seed=SEED_M, every=2):
    """Odd j: 1 + (h&1) substitutions at positions h_i mod k by (old + 1 + h_i mod 3) & 3."""
    x = np.array(kmers, dtype=np.uint64, copy=True)
    j = np.arange(len(x), dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64_np(np.uint64(seed) + j)
        nsub = np.uint64(1) + (h & np.uint64(1))
        sel = (j % np.uint64(every)) == np.uint64(every - 1)
        for r in range(2):
            hi = splitmix64_np(h + np.uint64(r + 1))
            p = hi % np.uint64(k)
            sh = np.uint64(2) * (np.uint64(k - 1) - p)
            old = (x >> sh) & np.uint64(3)
            new = (old + np.uint64(1) + (hi >> np.uint64(32)) % np.uint64(3)) & np.uint64(3)
            apply = sel & (nsub > np.uint64(r))
            x = np.where(apply, (x & ~(np.uint64(3) << sh)) | (new << sh), x)
    return x


def simulate_reads(genome: bytes, n_reads, length=150, sub_rate=0.01, seed=SEED_Q, n_rate=0.0005, revcomp_every=3):
    """Reads sampled from the genome: start = splitmix64(seed+j) mod (n-length), substitutions at
    `sub_rate`, a few N, every `revcomp_every`-th read reverse-complemented.  Deterministic."""
    n = len(genome)
    g = np.frombuffer(genome, dtype=np.uint8)
    rng = np.random.default_rng(seed & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        starts = (splitmix64_np(np.uint64(seed) + np.arange(n_reads, dtype=np.uint64)) % np.uint64(n - length)).astype(np.int64)
    comp = np.zeros(256, dtype=np.uint8)
    comp[:] = np.arange(256)
    for a, b in (b"AT", b"CG", b"GC", b"TA"):
        comp[a] = b
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for j, s0 in enumerate(starts):
        r = g[s0:s0 + length].copy()
        m = rng.random(length) < sub_rate
        r[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        r[rng.random(length) < n_rate] = ord("N")
        if revcomp_every and j % revcomp_every == revcomp_every - 1:
            r = comp[r[::-1]]
        reads.append(r.tobytes())
    return reads, starts


def unpack_kmer(x, k):
    return "".join("ACGT"[(int(x) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def write_fasta(path, genome: bytes, name="chr1", width=80):
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(genome), width):
            f.write(genome[i:i + width] + b"\n")

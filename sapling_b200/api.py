"""ctypes binding of libsapling_b200.so (include/sapling_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SaplingError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libsapling_b200.so")


_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

# name -> (restype, argtypes): every symbol include/sapling_b200.h declares
SYMBOLS = {
    "sapling_b200_open": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_uint]),
    "sapling_b200_create": (C.c_void_p, [C.c_char_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint]),
    "sapling_b200_create_with_model": (C.c_void_p, [C.c_char_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, _i64p, _i64p, _i32p, C.c_uint]),
    "sapling_b200_create_synthetic": (C.c_void_p, [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]),
    "sapling_b200_open_multi": (C.c_void_p, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_uint,
                                             C.c_uint64]),
    "sapling_b200_replicate": (C.c_int, [C.c_void_p, C.c_uint64]),
    "sapling_b200_num_devices": (C.c_int, [C.c_void_p]),
    "sapling_b200_close": (None, [C.c_void_p]),
    "sapling_b200_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)] + [C.POINTER(C.c_int)] * 7),
    "sapling_b200_genome": (C.c_void_p, [C.c_void_p]),
    "sapling_b200_num_chr": (C.c_size_t, [C.c_void_p]),
    "sapling_b200_chr": (C.c_uint64, [C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]),
    "sapling_b200_build_stats": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3),
    "sapling_b200_model": (C.c_int, [C.c_void_p, _i64p, _i64p]),
    "sapling_b200_rev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, _u32p]),
    "sapling_b200_sa_rank": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, _u32p]),
    "sapling_b200_write_sap": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sapling_b200_write_sa": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sapling_b200_device_bytes": (C.c_uint64, [C.c_void_p]),
    "sapling_b200_launch_count": (C.c_uint64, [C.c_void_p]),
    "sapling_b200_save_cache": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sapling_b200_open_cache": (C.c_void_p, [C.c_char_p, C.c_uint]),
    "sapling_b200_query_kernel": (C.c_char_p, [C.c_void_p, C.POINTER(C.c_int)]),
    "sapling_b200_query_kernel_for": (C.c_char_p, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "sapling_b200_query_partition_bits": (C.c_int, [C.c_void_p, C.c_size_t]),
    "sapling_b200_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "sapling_b200_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "sapling_b200_kmerize": (C.c_int64, [C.c_int, C.c_char_p]),
    "sapling_b200_kmerize_adjusted": (C.c_int64, [C.c_int, C.c_int, C.c_char_p]),
    "sapling_b200_query_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sapling_b200_query_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sapling_b200_query_batch_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "sapling_b200_query_batch_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "sapling_b200_query_batch_u32_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sapling_b200_query_str": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int64, C.c_size_t]),
    "sapling_b200_query_str_batch": (C.c_int, [C.c_void_p, C.c_char_p, _u64p, _u32p, C.c_void_p, _i64p, C.c_size_t, _i64p]),
    "sapling_b200_predict_batch": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, _u64p]),
    "sapling_b200_count_hits": (C.c_int, [C.c_void_p, _u32p, C.c_size_t, C.c_uint32, _u32p, _u32p]),
    "sapling_b200_seed_batch": (C.c_int, [C.c_void_p, C.c_char_p, _u64p, C.c_size_t, C.c_uint32, C.c_uint32, _i64p, _u32p, _u32p, _u32p]),
    "sapling_b200_seed_batch_compact": (C.c_int, [C.c_void_p, C.c_char_p, _u64p, C.c_size_t, C.c_uint32, C.c_uint32, _u32p, _u32p, _u8p, _u8p]),
    "sapling_b200_seed_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sapling_b200_oob_count": (C.c_uint64, [C.c_void_p]),
    "sapling_b200_sample_queries_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sapling_b200_verify_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]),
    "sapling_b200_count_probes_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_void_p]),
    "sapling_b200_check_sa": (C.c_int, [C.c_void_p, C.c_uint32] + [C.POINTER(C.c_uint64)] * 3),
    "sapling_b200_gather_bench": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "sapling_b200_gather_bench2": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "sapling_b200_last_error": (C.c_char_p, []),
    "sapling_b200_version": (C.c_char_p, []),
}


def lib():
    """Load libsapling_b200.so.  Fails loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise SaplingError(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C sapling_b200/csrc`")
        L = C.CDLL(p)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _err():
    return lib().sapling_b200_last_error().decode(errors="replace")


def _b(s):
    return s.encode() if isinstance(s, str) else s


def kmerize(k, s):
    """Sapling::kmerize (sapling_api.h:73-78)"""
    return int(lib().sapling_b200_kmerize(k, _b(s)))


def kmerize_adjusted(k, length, s):
    """Sapling::kmerizeAdjusted (sapling_api.h:83-90)"""
    return int(lib().sapling_b200_kmerize_adjusted(k, length, _b(s)))


def gather_bench(nbytes, n_loads, reps=3):
    g = C.c_double(0)
    if lib().sapling_b200_gather_bench(nbytes, n_loads, reps, C.byref(g)):
        raise SaplingError(_err())
    return g.value


def gather_bench2(nbytes, n_access, gran=32, chain=1, blocks_per_sm=8, reps=2):
    """1e9 random accesses/s of `gran` contiguous bytes over an nbytes buffer."""
    g = C.c_double(0)
    if lib().sapling_b200_gather_bench2(nbytes, n_access, gran, chain, blocks_per_sm, reps, C.byref(g)):
        raise SaplingError(_err())
    return g.value


class Sapling:
    """Mirror of the reference ``struct Sapling`` (sapling_api.h:17-679) over the C ABI.

    ``Sapling(refFn, saFn, sapFn, numBuckets=-1, maxMem=-1, k=-1, errorFn="")`` has the reference
    constructor's meaning (:492).  ``plQuery(s, kmer, length)`` is the reference query (:159);
    ``queryBatch(kmers)`` is the batched addition.
    """

    def __init__(self, refFn=None, saFn=None, sapFn=None, numBuckets=-1, maxMem=-1, k=-1, errorFn="",
                 flags=1, _handle=None):
        self._L = lib()
        if _handle is None:
            _handle = self._L.sapling_b200_open(_b(refFn), _b(saFn), _b(sapFn), numBuckets, maxMem, k,
                                                _b(errorFn or ""), flags)
        if not _handle:
            raise SaplingError(_err())
        self._h = C.c_void_p(_handle)
        n = C.c_uint64(0)
        v = [C.c_int(0) for _ in range(7)]
        self._L.sapling_b200_info(self._h, C.byref(n), *[C.byref(x) for x in v])
        self.n = n.value
        (self.k, self.buckets, self.maxOver, self.maxUnder, self.meanError, self.mostOver,
         self.mostUnder) = [x.value for x in v]

    # --- alternative constructors ---
    @classmethod
    def from_memory(cls, genome: bytes, sa=None, numBuckets=-1, maxMem=-1, k=-1, flags=1):
        L = lib()
        sap = None
        if sa is not None:
            sa = np.ascontiguousarray(sa, dtype=np.uint32)
            sap = sa.ctypes.data_as(C.c_void_p)
        return cls(_handle=L.sapling_b200_create(genome, len(genome), sap, numBuckets, maxMem, k, flags) or 0)

    @classmethod
    def from_model(cls, genome: bytes, sa, k, nb, xlist, ylist, five, flags=1):
        L = lib()
        sa = np.ascontiguousarray(sa, dtype=np.uint32)
        xl = np.ascontiguousarray(xlist, dtype=np.int64)
        yl = np.ascontiguousarray(ylist, dtype=np.int64)
        fv = np.ascontiguousarray(five, dtype=np.int32)
        h = L.sapling_b200_create_with_model(genome, len(genome), sa.ctypes.data_as(C.c_void_p), k, nb, xl, yl, fv, flags)
        return cls(_handle=h or 0)

    @classmethod
    def synthetic(cls, seed, n, numBuckets=-1, maxMem=-1, k=-1, keep_host_genome=False, flags=1):
        L = lib()
        return cls(_handle=L.sapling_b200_create_synthetic(seed, n, numBuckets, maxMem, k,
                                                           1 if keep_host_genome else 0, flags) or 0)

    @classmethod
    def open_multi(cls, refFn, saFn, sapFn, numBuckets=-1, maxMem=-1, k=-1, errorFn="", flags=1, gpu_mask=0):
        """The reference constructor, then replicas on the GPUs of gpu_mask (include/sapling_b200.h)."""
        L = lib()
        return cls(_handle=L.sapling_b200_open_multi(_b(refFn), _b(saFn), _b(sapFn), numBuckets, maxMem, k,
                                                     _b(errorFn or ""), flags, gpu_mask) or 0)

    @classmethod
    def from_cache(cls, path, flags=1):
        """Index restored from a private cache file written by save_cache."""
        L = lib()
        return cls(_handle=L.sapling_b200_open_cache(os.fsencode(path), flags) or 0)

    def save_cache(self, path):
        self._ck(self._L.sapling_b200_save_cache(self._h, os.fsencode(path)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.sapling_b200_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- members of the reference struct ---
    @property
    def five(self):
        return (self.maxOver, self.maxUnder, self.meanError, self.mostOver, self.mostUnder)

    @property
    def reference(self) -> bytes:
        p = self._L.sapling_b200_genome(self._h)
        if not p:
            raise SaplingError("host genome not kept for this index")
        if self.n < (1 << 31):
            return C.string_at(p, self.n)
        return bytes((C.c_char * self.n).from_address(p))  # string_at takes an int-sized length

    @property
    def chrEnds(self):
        out = []
        for i in range(self._L.sapling_b200_num_chr(self._h)):
            nm = C.c_char_p()
            pos = self._L.sapling_b200_chr(self._h, i, C.byref(nm))
            out.append((int(pos), nm.value.decode()))
        return sorted(out)

    @property
    def perfectPredictions(self):
        p = C.c_uint64(0)
        self._L.sapling_b200_build_stats(self._h, C.byref(p), None, None)
        return p.value

    def model(self):
        cnt = (1 << self.buckets) + 1
        x = np.empty(cnt, dtype=np.int64)
        y = np.empty(cnt, dtype=np.int64)
        self._ck(self._L.sapling_b200_model(self._h, x, y))
        return x, y

    def rev(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty(count, dtype=np.uint32)
        self._ck(self._L.sapling_b200_rev(self._h, first, count, out))
        return out

    def sa(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty(count, dtype=np.uint32)
        self._ck(self._L.sapling_b200_sa_rank(self._h, first, count, out))
        return out

    def device_bytes(self):
        return int(self._L.sapling_b200_device_bytes(self._h))

    def launch_count(self):
        return int(self._L.sapling_b200_launch_count(self._h))

    def partition_bits(self, nq):
        """Top k-mer bits a device batch of nq queries is partitioned by (0 = answered in the caller's order)."""
        return int(self._L.sapling_b200_query_partition_bits(self._h, nq))

    def profile(self, on=True):
        """Record CUDA events around the stages of every queryBatchDevice call (read them with stage_ms)."""
        self._ck(self._L.sapling_b200_profile(self._h, 1 if on else 0))

    def stage_ms(self):
        """(profiled calls, [histogram+scans, scatter, query kernel, un-permute] summed over them, ms); clears the list."""
        ms = (C.c_double * 4)()
        calls = self._L.sapling_b200_stage_ms(self._h, ms)
        if calls < 0:
            raise SaplingError(_err())
        return calls, [float(x) for x in ms]

    def query_kernel(self, nq=None):
        """(name of the CUDA kernel queryBatch launches for this index, resident blocks per SM it is compiled for);
        with nq: for a batch of that many queries (a partitioned batch runs the in-order kernel)."""
        b = C.c_int(0)
        if nq is None:
            name = self._L.sapling_b200_query_kernel(self._h, C.byref(b))
        else:
            name = self._L.sapling_b200_query_kernel_for(self._h, int(nq), C.byref(b))
        return (name or b"").decode(), b.value

    def _ck(self, rc):
        if rc != 0:
            raise SaplingError(_err())

    # --- methods of the reference struct ---
    def kmerize(self, s):
        return kmerize(self.k, s)

    def kmerizeAdjusted(self, length, s):
        return kmerize_adjusted(self.k, length, s)

    def queryPiecewiseLinear(self, kmers):
        kmers = np.ascontiguousarray(np.atleast_1d(kmers), dtype=np.uint64)
        out = np.empty(len(kmers), dtype=np.uint64)
        self._ck(self._L.sapling_b200_predict_batch(self._h, kmers, len(kmers), out))
        return out

    def plQuery(self, s, kmer, length):
        """long long plQuery(string s, long kmer, size_t length)   sapling_api.h:159"""
        s = _b(s)
        r = int(self._L.sapling_b200_query_str(self._h, s, len(s), int(kmer), int(length)))
        if r == -2:
            raise SaplingError(_err())
        return r

    def plQueryBatch(self, strings, kmers, lengths=None):
        strings = [_b(s) for s in strings]
        blob = b"".join(strings)
        slens = np.array([len(s) for s in strings], dtype=np.uint32)
        offs = np.zeros(len(strings), dtype=np.uint64)
        if len(strings) > 1:
            offs[1:] = np.cumsum(slens[:-1], dtype=np.uint64)
        km = np.ascontiguousarray(kmers, dtype=np.int64)
        out = np.empty(len(strings), dtype=np.int64)
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
            lp = lengths.ctypes.data_as(C.c_void_p)
        self._ck(self._L.sapling_b200_query_str_batch(self._h, blob, offs, slens, lp, km, len(strings), out))
        return out

    def queryBatch(self, kmers, out=None):
        """out[i] = plQuery(unpack(kmers[i]), kmers[i], k) for host arrays (numpy or pinned torch)."""
        ptr_in, nq = _host_ptr(kmers, 8)
        if out is None:
            out = np.empty(nq, dtype=np.int64)
        ptr_out, nq2 = _host_ptr(out, 8)
        assert nq2 >= nq
        self._ck(self._L.sapling_b200_query_batch(self._h, ptr_in, nq, ptr_out))
        return out

    def queryBatchU32(self, kmers, kmer_bytes=8, out=None, nq=None):
        """The narrow transfer format of the host path: `kmers` holds nq little-endian integers of kmer_bytes bytes each
        (a uint64 array for 8, a uint8 array of nq * kmer_bytes otherwise; numpy or pinned torch), the answers come
        back as uint32 positions with 0xFFFFFFFF for -1."""
        ptr_in, count = _host_ptr(kmers, 8 if kmer_bytes == 8 else 1)
        if nq is None:
            nq = count if kmer_bytes == 8 else count // kmer_bytes
        if out is None:
            out = np.empty(nq, dtype=np.uint32)
        ptr_out, nq2 = _host_ptr(out, 4)
        assert nq2 >= nq
        self._ck(self._L.sapling_b200_query_batch_u32(self._h, ptr_in, kmer_bytes, nq, ptr_out))
        return out

    def queryBatchBits(self, bits, kmer_bits, nq, out=None):
        """The densest upload of the host path: `bits` (uint8 array, numpy or pinned torch) is a little-endian bit stream
        with k-mer i in bits [i * kmer_bits, (i + 1) * kmer_bits), 2k <= kmer_bits <= 64 (`pack_kmer_bits` makes one);
        uint32 positions come back, 0xFFFFFFFF for -1."""
        ptr_in, count = _host_ptr(bits, 1)
        assert count * 8 >= nq * kmer_bits
        if out is None:
            out = np.empty(nq, dtype=np.uint32)
        ptr_out, nq2 = _host_ptr(out, 4)
        assert nq2 >= nq
        self._ck(self._L.sapling_b200_query_batch_bits(self._h, ptr_in, kmer_bits, nq, ptr_out))
        return out

    def queryBatchU32Device(self, d_kmers_ptr, nq, d_out_ptr, stream=0):
        self._ck(self._L.sapling_b200_query_batch_u32_dev(self._h, d_kmers_ptr, nq, d_out_ptr, stream))

    def replicate(self, gpu_mask):
        """Copies of the index on every other GPU whose bit is set (device to device); host batches then shard over them."""
        self._ck(self._L.sapling_b200_replicate(self._h, gpu_mask))
        return self.num_devices()

    def num_devices(self):
        return int(self._L.sapling_b200_num_devices(self._h))

    def queryBatchDevice(self, d_kmers_ptr, nq, d_out_ptr, stream=0):
        """Device-resident batch: raw device pointers (e.g. torch .data_ptr()), enqueued on `stream`."""
        self._ck(self._L.sapling_b200_query_batch_dev(self._h, d_kmers_ptr, nq, d_out_ptr, stream))

    def countHits(self, sa_pos, maxHits):
        sa_pos = np.ascontiguousarray(np.atleast_1d(sa_pos), dtype=np.uint32)
        left = np.empty(len(sa_pos), dtype=np.uint32)
        right = np.empty(len(sa_pos), dtype=np.uint32)
        self._ck(self._L.sapling_b200_count_hits(self._h, sa_pos, len(sa_pos), maxHits, left, right))
        return left, right

    def seedBatch(self, reads, num_seeds=7, max_hits=32):
        """Seed lookups of align.cpp seed_extend (:259-300) for a list of reads (bytes): returns (ref_pos, sa_pos, left,
        right), each of shape (n_reads, 2 strands, num_seeds); ref_pos is -1 where the seed has no verified hit."""
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        blob = b"".join(reads)
        m = len(reads) * 2 * num_seeds
        rp, sp = np.empty(m, np.int64), np.empty(m, np.uint32)
        lf, rt = np.empty(m, np.uint32), np.empty(m, np.uint32)
        self._ck(self._L.sapling_b200_seed_batch(self._h, blob, off, len(reads), num_seeds, max_hits, rp, sp, lf, rt))
        shp = (len(reads), 2, num_seeds)
        return rp.reshape(shp), sp.reshape(shp), lf.reshape(shp), rt.reshape(shp)

    def seedBatchCompact(self, reads, num_seeds=7, max_hits=32):
        """seedBatch in the 10-bytes-per-seed transfer format: ref_pos uint32 (0xFFFFFFFF = no verified hit), sa_pos uint32,
        left / right uint8; blocks of reads are pipelined through the GPU."""
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        blob = b"".join(reads)
        m = len(reads) * 2 * num_seeds
        rp, sp = np.empty(m, np.uint32), np.empty(m, np.uint32)
        lf, rt = np.empty(m, np.uint8), np.empty(m, np.uint8)
        self._ck(self._L.sapling_b200_seed_batch_compact(self._h, blob, off, len(reads), num_seeds, max_hits, rp, sp, lf, rt))
        shp = (len(reads), 2, num_seeds)
        return rp.reshape(shp), sp.reshape(shp), lf.reshape(shp), rt.reshape(shp)

    def seedBatchDevice(self, d_reads_ptr, d_off_ptr, n_reads, num_seeds, max_hits, d_ref_pos, d_sa_pos, d_left, d_right,
                        stream=0):
        self._ck(self._L.sapling_b200_seed_batch_dev(self._h, d_reads_ptr, d_off_ptr, n_reads, num_seeds, max_hits,
                                                     d_ref_pos, d_sa_pos, d_left, d_right, stream))

    def oob_count(self):
        return int(self._L.sapling_b200_oob_count(self._h))

    def write_sap(self, path):
        self._ck(self._L.sapling_b200_write_sap(self._h, _b(path)))

    def write_sa(self, path):
        self._ck(self._L.sapling_b200_write_sa(self._h, _b(path)))

    def check_sa(self, max_chars=1 << 20):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._ck(self._L.sapling_b200_check_sa(self._h, max_chars, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sample_queries_device(self, seed, mut_seed, first, nq, d_kmers_ptr, stream=0):
        self._ck(self._L.sapling_b200_sample_queries_dev(self._h, seed, mut_seed, first, nq, d_kmers_ptr, stream))

    def verify_device(self, d_kmers_ptr, d_out_ptr, nq, stream=0):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._L.sapling_b200_verify_dev(self._h, d_kmers_ptr, d_out_ptr, nq, C.byref(a), C.byref(b), stream))
        return a.value, b.value

    def count_probes_device(self, d_kmers_ptr, nq, stream=0):
        """Total getLcp calls the reference's plQuery makes for these queries (SURVEY 8d's P x nq)."""
        a = C.c_uint64(0)
        self._ck(self._L.sapling_b200_count_probes_dev(self._h, d_kmers_ptr, nq, C.byref(a), stream))
        return a.value


def pack_kmer_bits(kmers, kmer_bits):
    """k-mers (integers < 2^kmer_bits) -> the little-endian bit stream `queryBatchBits` uploads: k-mer i in bits
    [i * kmer_bits, (i + 1) * kmer_bits).  Eight k-mers make kmer_bits whole bytes, so the stream is built in groups of 8."""
    x = np.ascontiguousarray(kmers, dtype=np.uint64)
    n = len(x)
    g = np.zeros(((n + 7) // 8) * 8, dtype=np.uint64)
    g[:n] = x
    g = g.reshape(-1, 8)
    nw = (8 * kmer_bits + 63) // 64
    W = np.zeros((g.shape[0], nw + 1), dtype=np.uint64)
    for j in range(8):
        w, sh = divmod(j * kmer_bits, 64)
        W[:, w] |= g[:, j] << np.uint64(sh)
        if sh and sh + kmer_bits > 64:
            W[:, w + 1] |= g[:, j] >> np.uint64(64 - sh)
    stream = W.view(np.uint8).reshape(g.shape[0], (nw + 1) * 8)[:, :kmer_bits].reshape(-1)
    return np.ascontiguousarray(stream[: (n * kmer_bits + 7) // 8])


def _host_ptr(a, itemsize):
    """(void*, element count) of a numpy array or a CPU torch tensor of `itemsize`-byte elements."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"] and a.dtype.itemsize == itemsize
        return C.c_void_p(a.ctypes.data), a.size
    # torch tensor
    assert a.is_contiguous() and a.element_size() == itemsize and a.device.type == "cpu"
    return C.c_void_p(a.data_ptr()), a.numel()

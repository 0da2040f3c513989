// partition.cu -- batch partitioning around the query kernel.
//
// Why.  A k-mer's bucket (top nb bits) and its suffix-array rank are both monotone in the k-mer value, so the top
// bits of a query say which SLICE of the model table and of the suffix array (or rank lines) it will touch.  An
// unordered batch gathers from the whole multi-GB index at once: every model read and every suffix-array read is a
// TLB miss plus a 128-byte DRAM line fill that no other query reuses (profiles/r1_experiments.md sections 1-2).  If
// the batch is first partitioned by the top `pbits` bits of the k-mer and the query kernel then walks the partitioned
// array in order, the ~150 k queries in flight at any moment all fall into one or two slices: the model slice and
// (for a fine enough partition) the suffix-array slice are L2 resident, a DRAM line is filled at most once per batch
// and shared by all queries that need it, and the SMs touch a handful of 2 MB pages instead of thousands.
// The price is three streaming passes (42 bytes per query); the reference's per-query arithmetic is untouched --
// the same Replay code answers every query, only the order in which queries are answered changes.
//
//   A  part_hist_kernel     chunk c (kPartChunk queries) x bin -> count            cnt[c][bin]  (chunk-major: every
//                           pass reads or writes whole rows; a bin-major table cost 1024 strided sectors per chunk)
//   S  part_col*_kernel     exclusive scan down every column (over chunks), then over bins     off[c][bin], bin_start[bin]
//   B  part_scatter_staged_kernel  k-mer -> part_kmer[bin_start + off + local rank], its index inside the chunk (slot) with it
//   Q  query kernel (query.cu) over part_kmer in order; result word = slot << 48 | answer
//   U  part_unpermute_kernel  chunk c gathers its answers bin by bin into shared memory, writes out[] coalesced
#include "partition.cuh"

#include <mutex>

namespace sb {

namespace {

constexpr int kHistThreads = 512;
constexpr int kUnpermThreads = 512;  // the warp-per-run kernel: two blocks per SM at up to 64 registers, 8 answers in flight per lane

__device__ __forceinline__ uint32_t bin_of(uint64_t x, int pshift, uint32_t nbins) {
  const uint64_t b = x >> pshift;
  return b < (uint64_t)nbins ? (uint32_t)b : nbins - 1;  // a k-mer with bits above 2k set: still a valid array slot
}

// A: per-chunk histogram.  cnt is chunk-major: row c holds the nbins counts of chunk c (row nchunks: the totals).
__global__ void __launch_bounds__(kHistThreads)
part_hist_kernel(const uint64_t* __restrict__ kmers, size_t nq, int pshift, uint32_t nbins, size_t nchunks,
                 uint32_t* __restrict__ cnt) {
  extern __shared__ uint32_t sh[];
  const size_t c = blockIdx.x;
  for (uint32_t b = threadIdx.x; b < nbins; b += kHistThreads) sh[b] = 0;
  __syncthreads();
  const size_t base = c * kPartChunk;
  const uint32_t m = (uint32_t)(nq - base < kPartChunk ? nq - base : kPartChunk);
  for (uint32_t i = threadIdx.x; i < m; i += kHistThreads) atomicAdd(&sh[bin_of(__ldg(kmers + base + i), pshift, nbins)], 1u);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nbins; b += kHistThreads) cnt[c * nbins + b] = sh[b];
}

// S1: exclusive scan down every column of cnt (over the chunks of one bin), in three small passes so that every access is
// a coalesced row piece: the rows are cut into kScanSegs segments; (a) per-segment column sums, (b) a serial scan of the
// kScanSegs sums per column (totals -> row nchunks), (c) each segment rewrites its rows as running sums.
constexpr int kScanSegs = 64;
__global__ void __launch_bounds__(128)
part_colsum_kernel(const uint32_t* __restrict__ cnt, size_t nchunks, uint32_t nbins, size_t rows_per_seg,
                   uint32_t* __restrict__ segsum) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t r0 = (size_t)blockIdx.y * rows_per_seg;
  const size_t r1 = r0 + rows_per_seg < nchunks ? r0 + rows_per_seg : nchunks;
  uint32_t sum = 0;
  for (size_t r = r0; r < r1; r++) sum += __ldg(cnt + r * nbins + col);
  segsum[(size_t)blockIdx.y * nbins + col] = sum;
}
__global__ void __launch_bounds__(128)
part_segscan_kernel(uint32_t* __restrict__ segsum, uint32_t nbins, uint32_t* __restrict__ totals) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v[kScanSegs];  // all loads in flight before the first store (a load-store-load chain cost 16 us per launch)
#pragma unroll
  for (int sg = 0; sg < kScanSegs; sg++) v[sg] = __ldg(segsum + (size_t)sg * nbins + col);
  uint32_t run = 0;
#pragma unroll
  for (int sg = 0; sg < kScanSegs; sg++) {
    segsum[(size_t)sg * nbins + col] = run;
    run += v[sg];
  }
  totals[col] = run;
}
__global__ void __launch_bounds__(128)
part_colscan_kernel(uint32_t* __restrict__ cnt, size_t nchunks, uint32_t nbins, size_t rows_per_seg,
                    const uint32_t* __restrict__ segsum) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t r0 = (size_t)blockIdx.y * rows_per_seg;
  const size_t r1 = r0 + rows_per_seg < nchunks ? r0 + rows_per_seg : nchunks;
  uint32_t run = segsum[(size_t)blockIdx.y * nbins + col];
  for (size_t r = r0; r < r1; r++) {
    const uint32_t v = cnt[r * nbins + col];
    cnt[r * nbins + col] = run;
    run += v;
  }
}

// S2: one block: bin_start = exclusive scan of the row totals (nbins <= 2048)
__global__ void __launch_bounds__(1024)
part_scan_bins_kernel(const uint32_t* __restrict__ cnt, size_t nchunks, uint32_t nbins, uint32_t* __restrict__ bin_start,
                      unsigned long long* __restrict__ tiles) {
  __shared__ uint32_t a[2048];
  if (threadIdx.x == 0) *tiles = 0;  // the query kernel's in-order tile counter (query.cu QueryCursor)
  for (uint32_t b = threadIdx.x; b < 2048; b += 1024) a[b] = b < nbins ? cnt[nchunks * nbins + b] : 0u;
  __syncthreads();
  for (int d = 1; d < 2048; d <<= 1) {
    uint32_t v[2];
    for (int j = 0; j < 2; j++) {
      const uint32_t b = threadIdx.x + 1024u * j;
      v[j] = b >= (unsigned)d ? a[b - d] : 0u;
    }
    __syncthreads();
    for (int j = 0; j < 2; j++) a[threadIdx.x + 1024u * j] += v[j];
    __syncthreads();
  }
  for (uint32_t b = threadIdx.x; b < nbins; b += 1024) bin_start[b] = b ? a[b - 1] : 0u;
  if (threadIdx.x == 0) bin_start[nbins] = a[nbins - 1];
}

// Block-wide exclusive scan of 2048 counters in shared memory; thread t owns the kPer = 2048 / kThreads consecutive
// counters from t * kPer.  On return a[b] = sum of the counts before b; own[j] holds the thread's own exclusive starts.
template <int kThreads>
__device__ __forceinline__ void scan_2048(uint32_t* a, uint32_t* wsum, uint32_t (&own)[2048 / kThreads]) {
  constexpr int kPer = 2048 / kThreads;
  const uint32_t b0 = (uint32_t)kPer * threadIdx.x;
  uint32_t c[kPer];
  uint32_t total = 0;
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    c[j] = a[b0 + j];
    total += c[j];
  }
  uint32_t incl = total;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((threadIdx.x & 31u) >= (unsigned)d) incl += v;
  }
  if ((threadIdx.x & 31u) == 31u) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = threadIdx.x < kThreads / 32 ? wsum[threadIdx.x] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
      if (threadIdx.x >= (unsigned)d) w += v;
    }
    wsum[threadIdx.x] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t run = ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u) + incl - total;
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    own[j] = run;
    a[b0 + j] = run;
    run += c[j];
  }
}

// B: the scatter, staged through shared memory.  A direct scatter sends every warp store to 32 different lines (measured:
// 1.0 ms per 50 M queries).  Here a block first sorts its chunk by bin inside shared memory and then writes the sorted
// chunk out in order: consecutive threads store consecutive addresses for as long as the bin lasts (chunk / bins queries
// on average).  The order of a chunk's queries inside one bin is whatever the shared-memory atomics make it: the slot
// travels with the query, so the un-permute pass does not care.
// The (chunk, bin) counts are already known -- pass A wrote them, the scan turned them into offsets -- so the block
// does not count again: it requests its column of the offset table together with its k-mers (each thread keeps 16 of
// them in registers), scans the run lengths into run starts, and then ONE shared-memory atomic per k-mer hands out the
// k-mer's place in the sorted chunk.  (The version measured in gpurun r2d counted with returning atomics, kept the
// ranks in shared memory and re-read the run starts: 9 shared-memory operations per k-mer against 6 here.)
constexpr int kScatterThreads = kPartChunk / 16;  // 16 k-mers per thread in registers
constexpr int kScatterPer = kPartChunk / kScatterThreads;
// (Measured and dropped, gpurun s22: the write-out as one bulk copy per run -- cp.async.bulk shared -> global, runs padded in
// the staging buffer to the parity of their global index so that both sides are 16-byte aligned together -- takes 128 of a
// thread's ~740 instructions per chunk away and changes nothing: 1.055 against 1.057 ms.  The pass waits for its k-mers
// (29 % of the stall samples) and at its barriers: one staging buffer per SM means load, sort and write-out of a chunk
// take turns, and a second buffer does not fit.)
// Persistent: one block per SM (the staging buffer takes most of its shared memory) walks chunks blockIdx, blockIdx +
// gridDim, ...  With a block per chunk the phases of a chunk ran one after the other -- load, scan, sort, write out -- and
// the SM sat idle through every load (ncu r2u: 81 % of the stalls long-scoreboard, 49 % of the copy bandwidth).  Here the
// k-mers and the offset-table rows of the NEXT chunk are requested as soon as the current chunk has been sorted into shared
// memory (its k-mer registers are dead by then, so they are simply reused) and arrive while the current chunk is written
// out.
template <bool kSlotInKmer>
__global__ void __launch_bounds__(kScatterThreads)
part_scatter_staged_kernel(const uint64_t* __restrict__ kmers, size_t nq, int pshift, uint32_t nbins, size_t nchunks,
                           const uint32_t* __restrict__ off, const uint32_t* __restrict__ bin_start,
                           uint64_t* __restrict__ part_kmer, uint16_t* __restrict__ part_slot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);                         // [kPartChunk] k-mers sorted by bin
  uint16_t* ss = reinterpret_cast<uint16_t*>(sk + kPartChunk);                  // [kPartChunk] their slots
  uint32_t* lstart = reinterpret_cast<uint32_t*>(ss + kPartChunk);              // [2048] next free sorted index of the bin
  uint32_t* gdelta = lstart + 2048;                                             // [2048] global position - sorted index
  __shared__ uint32_t wsum[32];
  constexpr int kBins = 2048 / kScatterThreads;  // bins per thread
  uint64_t x[kScatterPer];
  uint32_t r0[kBins], r1[kBins], bs[kBins];
  auto request = [&](size_t c) {  // the k-mers of chunk c and this thread's piece of rows c, c + 1 of the offset table
    const size_t base = c * kPartChunk;
    const uint32_t m = (uint32_t)(nq - base < kPartChunk ? nq - base : kPartChunk);
#pragma unroll
    for (int j = 0; j < kScatterPer; j++) {
      const uint32_t i = threadIdx.x + (uint32_t)j * kScatterThreads;
      x[j] = i < m ? __ldcs(kmers + base + i) : 0ull;
    }
#pragma unroll
    for (int j = 0; j < kBins; j++) {
      const uint32_t b = (uint32_t)kBins * threadIdx.x + j;
      r0[j] = r1[j] = bs[j] = 0;
      if (b < nbins) {
        const uint32_t* row = off + c * nbins + b;  // coalesced
        r0[j] = __ldg(row);
        r1[j] = __ldg(row + nbins);
        bs[j] = __ldg(bin_start + b);
      }
    }
  };
  size_t c = blockIdx.x;
  if (c < nchunks) request(c);
  for (; c < nchunks; c += gridDim.x) {
    const size_t base = c * kPartChunk;
    const uint32_t m = (uint32_t)(nq - base < kPartChunk ? nq - base : kPartChunk);
    uint32_t gpos[kBins];
#pragma unroll
    for (int j = 0; j < kBins; j++) {
      gpos[j] = bs[j] + r0[j];
      lstart[(uint32_t)kBins * threadIdx.x + j] = r1[j] - r0[j];
    }
    __syncthreads();
    {
      uint32_t own[kBins];
      scan_2048<kScatterThreads>(lstart, wsum, own);  // run lengths -> first sorted index of every run
#pragma unroll
      for (int j = 0; j < kBins; j++) gdelta[(uint32_t)kBins * threadIdx.x + j] = gpos[j] - own[j];
    }
    __syncthreads();
    // four atomics in flight per thread, then their four stores (more would spill: the 16 k-mers hold 32 registers)
#pragma unroll
    for (int h = 0; h < kScatterPer; h += 4) {
      uint32_t p[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t i = threadIdx.x + (uint32_t)(h + j) * kScatterThreads;
        p[j] = i < m ? atomicAdd(&lstart[bin_of(x[h + j] & (kSlotInKmer ? kSlotKmerMask : ~0ull), pshift, nbins)], 1u) : 0u;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t i = threadIdx.x + (uint32_t)(h + j) * kScatterThreads;
        if (i < m) {
          sk[p[j]] = kSlotInKmer ? ((x[h + j] & kSlotKmerMask) | ((uint64_t)i << kSlotShift)) : x[h + j];
          if (!kSlotInKmer) ss[p[j]] = (uint16_t)i;
        }
      }
    }
    if (c + gridDim.x < nchunks) request(c + gridDim.x);  // in flight during the write-out below
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += kScatterThreads) {
      const uint64_t v = sk[i];
      const uint32_t g = gdelta[bin_of(kSlotInKmer ? (v & kSlotKmerMask) : v, pshift, nbins)] + i;
      part_kmer[g] = v;
      if (!kSlotInKmer) part_slot[g] = ss[i];
    }
    __syncthreads();  // the staging buffer and the tables are rewritten by the next chunk
  }
}

// U: chunk c collects its answers.  Warp w walks bins w, w+32, ...; the run (bin, c) is read coalesced by the lanes.
// Out = long long (the reference's return type, -1 for "not found") or uint32_t (0xFFFFFFFF for -1: the host path's
// download format).
template <typename Out>
__device__ __forceinline__ void store_answer(Out* p, uint32_t v);
template <>
__device__ __forceinline__ void store_answer<long long>(long long* p, uint32_t v) {
  __stcs(p, v == 0xFFFFFFFFu ? -1ll : (long long)v);
}
template <>
__device__ __forceinline__ void store_answer<uint32_t>(uint32_t* p, uint32_t v) { __stcs(p, v); }

// (Measured and dropped, gpurun s27: the whole gather of a chunk as cp.async copies into a 128 KB staging buffer -- every
// answer of the chunk in flight at once, permuted inside shared memory afterwards, one persistent block per SM -- 1.038
// against 1.031 ms at c3; likewise 512 threads with 16 answers in flight per lane, gpurun s14: slower.  With the scatter's
// bulk-copy write-out (partition.cu above) that makes three designs of these two passes at the same 3.9 TB/s: what binds
// them is the memory side of 256-byte runs in 512 streams, not how the SM issues them.)
// Both kernels below request the answers of a step together, before the first shared-memory store that depends on one: an
// SM issues in order, so a loop of "load, store what was loaded" waits out one DRAM round trip per answer.
template <typename Out>
__global__ void __launch_bounds__(kUnpermThreads, 2)
part_unpermute_kernel(const long long* __restrict__ res, size_t nq, uint32_t nbins, size_t nchunks,
                      const uint32_t* __restrict__ off, const uint32_t* __restrict__ bin_start, Out* __restrict__ out) {
  // answers staged as 32 bits (a position < n <= 2^32 - 16, or 0xFFFFFFFF for -1): 64 KB per chunk, so that two blocks
  // share an SM and one block's loads overlap the other's stores
  extern __shared__ uint32_t buf32[];
  const size_t c = blockIdx.x;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x / 32u;
  constexpr uint32_t kWarps = kUnpermThreads / 32;
  constexpr int kAhead = 8;  // answers in flight per lane
  // warp w walks the runs of bins w, w + kWarps, ...: 32 consecutive answers per step, kAhead steps requested at once
  for (uint32_t b = warp; b < nbins; b += kWarps) {
    const uint32_t* row = off + c * nbins + b;
    const uint32_t s = __ldg(bin_start + b);
    const uint32_t a = s + __ldg(row), e = s + __ldg(row + nbins);
    for (uint32_t p0 = a + lane; p0 < e; p0 += 32u * kAhead) {
      unsigned long long v[kAhead];
#pragma unroll
      for (int t = 0; t < kAhead; t++) {
        const uint32_t p = p0 + 32u * (uint32_t)t;
        v[t] = p < e ? (unsigned long long)__ldcs(res + p) : 0ull;
      }
#pragma unroll
      for (int t = 0; t < kAhead; t++)
        if (p0 + 32u * (uint32_t)t < e) buf32[v[t] >> 48] = (uint32_t)v[t];  // low 32 bits: -1 reads back as 0xFFFFFFFF
    }
  }
  __syncthreads();
  const size_t base = c * kPartChunk;
  const uint32_t m = (uint32_t)(nq - base < kPartChunk ? nq - base : kPartChunk);
  for (uint32_t i = threadIdx.x; i < m; i += kUnpermThreads) store_answer<Out>(out + base + i, buf32[i]);
}

// U': lane GROUPS per run.  Above, a warp walks whole (bin, chunk) runs, which leaves most lanes idle once runs are shorter
// than a warp (1024 bins: 16 answers per run).  Here kGroup lanes (the power of two at or above the mean run length) own
// a run; per step a lane takes kBatch runs: their bounds are requested together, then its first answer of every run --
// kBatch loads in flight -- and only then the stores; what a run has beyond kGroup answers follows in a tail loop.
// 1024 threads, two blocks per SM (32 registers): measured against 512 threads with 16 loads in flight per lane (48
// registers, gpurun s14): 1.23 against 1.49 ms at c3 -- the pass lives on resident warps, every run is a new DRAM page.
constexpr int kGroupThreads = 1024;
template <int kGroup, typename Out>
__global__ void __launch_bounds__(kGroupThreads, 2)
part_unpermute_group_kernel(const long long* __restrict__ res, size_t nq, uint32_t nbins, size_t nchunks,
                            const uint32_t* __restrict__ off, const uint32_t* __restrict__ bin_start,
                            Out* __restrict__ out) {
  extern __shared__ uint32_t buf32[];  // [kPartChunk] answers in the caller's order, 32 bits each (see part_unpermute_kernel)
  constexpr uint32_t kGroups = kGroupThreads / kGroup;
  constexpr int kBatch = 4;
  const size_t c = blockIdx.x;
  const uint32_t g = threadIdx.x / kGroup, l = threadIdx.x % kGroup;
  for (uint32_t b0 = g; b0 < nbins; b0 += kGroups * kBatch) {
    uint32_t lo[kBatch], hi[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; j++) {
      const uint32_t b = b0 + (uint32_t)j * kGroups;
      lo[j] = hi[j] = 0;
      if (b < nbins) {
        const uint32_t* row = off + c * nbins + b;
        const uint32_t s = __ldg(bin_start + b);
        lo[j] = s + __ldg(row) + l;
        hi[j] = s + __ldg(row + nbins);
      }
    }
    unsigned long long v[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; j++) v[j] = lo[j] < hi[j] ? (unsigned long long)__ldcs(res + lo[j]) : 0ull;
#pragma unroll
    for (int j = 0; j < kBatch; j++)
      if (lo[j] < hi[j]) buf32[v[j] >> 48] = (uint32_t)v[j];
#pragma unroll
    for (int j = 0; j < kBatch; j++) {
      for (uint32_t p = lo[j] + kGroup; p < hi[j]; p += kGroup) {  // the rest of a long run
        const unsigned long long w = (unsigned long long)__ldcs(res + p);
        buf32[w >> 48] = (uint32_t)w;
      }
    }
  }
  __syncthreads();
  const size_t base = c * kPartChunk;
  const uint32_t m = (uint32_t)(nq - base < kPartChunk ? nq - base : kPartChunk);
  for (uint32_t i = threadIdx.x; i < m; i += kGroupThreads) store_answer<Out>(out + base + i, buf32[i]);
}

}  // namespace

constexpr size_t kScatterSmem = (size_t)kPartChunk * 10 + 2 * 2048 * 4;
constexpr size_t kUnpermSmem = (size_t)kPartChunk * sizeof(uint32_t);

size_t partition_workspace_bytes(size_t nq, int pbits) {
  const size_t nchunks = (nq + kPartChunk - 1) / kPartChunk;
  const size_t nbins = (size_t)1 << pbits;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return up(nq * 8) + up(nq * 8) + up(nq * 2) + up(nbins * (nchunks + 1) * 4) + up((nbins + 1) * 4) + 256 +
         up((size_t)kScanSegs * nbins * 4);
}

// Opt-in to the large dynamic shared memory of the scatter and un-permute kernels.  Function attributes belong to the
// device (context) they were set on, and one process may hold indexes on several GPUs: once per device.
static int set_kernel_attributes() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  SB_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
#define SB_SMEM(kernel, bytes)                                                                              \
  SB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
  SB_SMEM(part_scatter_staged_kernel<false>, kScatterSmem);
  SB_SMEM(part_scatter_staged_kernel<true>, kScatterSmem);
  SB_SMEM(part_unpermute_kernel<long long>, kUnpermSmem);
  SB_SMEM(part_unpermute_kernel<uint32_t>, kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<8, long long>), kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<16, long long>), kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<32, long long>), kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<8, uint32_t>), kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<16, uint32_t>), kUnpermSmem);
  SB_SMEM((part_unpermute_group_kernel<32, uint32_t>), kUnpermSmem);
#undef SB_SMEM
  cudaGetLastError();
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
}

template <typename Out>
static void launch_unpermute(const long long* res, size_t nq, uint32_t nbins, size_t nchunks, const uint32_t* cnt,
                             const uint32_t* bin_start, Out* out, int pbits, cudaStream_t st) {
  // lane groups per run, the group at the mean run length; a whole warp per run from 64 answers per run
  const uint32_t mean_run = kPartChunk >> pbits;
  const unsigned g = (unsigned)nchunks;
  if (mean_run >= 64)
    part_unpermute_kernel<Out><<<g, kUnpermThreads, kUnpermSmem, st>>>(res, nq, nbins, nchunks, cnt, bin_start, out);
  else if (mean_run >= 32)
    part_unpermute_group_kernel<32, Out><<<g, kGroupThreads, kUnpermSmem, st>>>(res, nq, nbins, nchunks, cnt, bin_start, out);
  else if (mean_run >= 16)
    part_unpermute_group_kernel<16, Out><<<g, kGroupThreads, kUnpermSmem, st>>>(res, nq, nbins, nchunks, cnt, bin_start, out);
  else
    part_unpermute_group_kernel<8, Out><<<g, kGroupThreads, kUnpermSmem, st>>>(res, nq, nbins, nchunks, cnt, bin_start, out);
}

int launch_partitioned_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                             void* ws, int pbits, cudaStream_t st, cudaEvent_t* ev) {
  if (nq == 0) return 0;
  if (pbits < 1 || pbits > kPartMaxBits || pbits > 2 * ix.k || nq >= (1ull << 32)) {
    set_error("launch_partitioned_query: pbits=%d nq=%zu out of range", pbits, nq);
    return -1;
  }
  if (set_kernel_attributes()) return -1;
  const size_t nchunks = (nq + kPartChunk - 1) / kPartChunk;
  const uint32_t nbins = 1u << pbits;
  const int pshift = 2 * ix.k - pbits;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  char* p = static_cast<char*>(ws);
  uint64_t* part_kmer = reinterpret_cast<uint64_t*>(p); p += up(nq * 8);
  long long* res = reinterpret_cast<long long*>(p); p += up(nq * 8);
  uint16_t* part_slot = reinterpret_cast<uint16_t*>(p); p += up(nq * 2);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(p); p += up((size_t)nbins * (nchunks + 1) * 4);
  uint32_t* bin_start = reinterpret_cast<uint32_t*>(p); p += up((size_t)(nbins + 1) * 4);
  unsigned long long* tiles = reinterpret_cast<unsigned long long*>(p); p += 256;
  uint32_t* segsum = reinterpret_cast<uint32_t*>(p);

  if (ev) cudaEventRecord(ev[0], st);
  part_hist_kernel<<<(unsigned)nchunks, kHistThreads, nbins * 4, st>>>(d_kmers, nq, pshift, nbins, nchunks, cnt);
  {
    const unsigned cw = nbins < 128u ? nbins : 128u;  // columns per block (nbins is a power of two)
    const size_t rows_per_seg = (nchunks + kScanSegs - 1) / kScanSegs;
    const dim3 grid(nbins / cw, kScanSegs);
    part_colsum_kernel<<<grid, cw, 0, st>>>(cnt, nchunks, nbins, rows_per_seg, segsum);
    part_segscan_kernel<<<nbins / cw, cw, 0, st>>>(segsum, nbins, cnt + nchunks * nbins);
    part_colscan_kernel<<<grid, cw, 0, st>>>(cnt, nchunks, nbins, rows_per_seg, segsum);
  }
  part_scan_bins_kernel<<<1, 1024, 0, st>>>(cnt, nchunks, nbins, bin_start, tiles);
  if (ev) cudaEventRecord(ev[1], st);
  // the slot of a query rides in bits 50-63 of its k-mer word when the k-mer leaves room (k <= 25): one store and one
  // load per query instead of two of each
  const bool slot_in_kmer = 2 * ix.k + 14 <= 64;
  const unsigned sgrid = (unsigned)(nchunks < 148 ? nchunks : 148);  // persistent: one block per SM
  if (slot_in_kmer)
    part_scatter_staged_kernel<true><<<sgrid, kScatterThreads, kScatterSmem, st>>>(
        d_kmers, nq, pshift, nbins, nchunks, cnt, bin_start, part_kmer, part_slot);
  else
    part_scatter_staged_kernel<false><<<sgrid, kScatterThreads, kScatterSmem, st>>>(
        d_kmers, nq, pshift, nbins, nchunks, cnt, bin_start, part_kmer, part_slot);
  SB_CUDA_CHECK(cudaGetLastError());
  if (ev) cudaEventRecord(ev[2], st);
  if (launch_kmer_query_ordered(ix, part_kmer, nq, res, slot_in_kmer ? slot_in_kmer_tag() : part_slot, tiles, st))
    return -1;
  if (ev) cudaEventRecord(ev[3], st);
  if (d_out32) launch_unpermute<uint32_t>(res, nq, nbins, nchunks, cnt, bin_start, d_out32, pbits, st);
  else launch_unpermute<long long>(res, nq, nbins, nchunks, cnt, bin_start, d_out, pbits, st);
  if (ev) cudaEventRecord(ev[4], st);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sb

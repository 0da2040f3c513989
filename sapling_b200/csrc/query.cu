// query.cu -- query kernels and their launchers.
//
// Hot path: kmer_query_kernel.  One thread per query, grid-stride over the batch so that the
// k-mer reads and result writes are coalesced; every dependent access (model checkpoint pair,
// SA entry, packed-genome window) is a read-only 32-byte-sector gather (ld.global.nc).  The
// kernel is bound by random-sector HBM/L2 throughput and by the length of the dependent chain, so
// it runs at full occupancy (see DESIGN.md for the roofline and profiles/ for ncu evidence).
#include "build.cuh"
#include "common.cuh"
#include "query.cuh"

namespace sb {

namespace {

constexpr int kQueryThreads = 256;

// Which query a lane answers next.
//   tiles == nullptr: grid-stride over the batch (k-mer reads and result writes coalesced; the order in which blocks
//     reach which part of the batch does not matter for an unordered batch).
//   tiles != nullptr: the batch is partitioned (partition.cu) and must be WALKED IN ORDER for the partition to pay: a
//     warp takes the next 32 queries from a global counter, so the queries in flight on the whole GPU are always the
//     ~300 k that follow each other in the partitioned array -- one or two slices of the index.  With a static
//     grid-stride schedule the blocks drift apart by several slices and the slices fall out of L2 and of TLB reach
//     again (ncu, profiles/r1y: 273 DRAM bytes per query where the partitioned order needs ~150).  The counter for the
//     next tile is bumped before the current tile is answered, so its round trip never sits on the critical path.
struct QueryCursor {
  size_t i, stride;
  unsigned long long next_tile;
  unsigned long long* tiles;
  __device__ __forceinline__ void init(unsigned long long* t) {
    tiles = t;
    if (tiles) {
      unsigned long long a = 0, b = 0;
      if ((threadIdx.x & 31u) == 0) {
        a = atomicAdd(tiles, 32ull);
        b = atomicAdd(tiles, 32ull);
      }
      i = (size_t)__shfl_sync(0xffffffffu, a, 0) + (threadIdx.x & 31u);
      next_tile = __shfl_sync(0xffffffffu, b, 0);
      stride = 0;
    } else {
      stride = (size_t)gridDim.x * blockDim.x;
      i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
      next_tile = 0;
    }
  }
  // warp-uniform in tile mode (a tile is in range when its first query is)
  __device__ __forceinline__ bool more(size_t nq) const { return tiles ? (i - (threadIdx.x & 31u)) < nq : i < nq; }
  __device__ __forceinline__ void next() {
    if (tiles) {
      unsigned long long b = 0;
      if ((threadIdx.x & 31u) == 0) b = atomicAdd(tiles, 32ull);
      i = (size_t)next_tile + (threadIdx.x & 31u);
      next_tile = __shfl_sync(0xffffffffu, b, 0);
    } else {
      i += stride;
    }
  }
};

template <int kMinBlocks>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t x = (ix.hints & HINT_IO_STREAM) ? __ldcs(kmers + i) : __ldg(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    const long long r = pl_query<false>(ix, q, x);
    if (ix.hints & HINT_IO_STREAM) __stcs(out + i, r);
    else out[i] = r;
  }
}

// Software-pipelined variant (narrow model layout only).  While a thread replays query t it already has in
// flight: the first suffix-array entry of query t+1, the model checkpoints of query t+2 and the k-mer of
// query t+3, so three of the dependent DRAM round trips of a query (k-mer -> model -> rev[predicted]) are
// overlapped with the probe chain of the previous queries instead of heading it.
template <int kMinBlocks>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_pipelined_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq,
                            long long* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= nq) return;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  const size_t last = nq - 1;
  auto kmer_at = [&](size_t i) { return __ldcs(kmers + (i < last ? i : last)); };

  // prologue: fill the pipeline
  uint64_t x0 = kmer_at(i0);
  uint64_t x1 = kmer_at(i0 + stride);
  uint64_t x2 = kmer_at(i0 + 2 * stride);
  uint64_t pred0 = clamp_prediction(ix, narrow_finish(ix, x0, narrow_load(ix, x0, pol.model), pol.model));
  uint32_t idx0 = ld_u32_pol(ix.sa + pred0, pol.sa);
  NarrowPair m1 = narrow_load(ix, x1, pol.model);

  for (size_t i = i0; i < nq; i += stride) {
    // stage C: k-mer of query t+3
    const uint64_t x3 = kmer_at(i + 3 * stride);
    // stage B: model checkpoints of query t+2 (x2 was loaded one iteration ago)
    const NarrowPair m2 = narrow_load(ix, x2, pol.model);
    // stage A: prediction of query t+1 (checkpoints loaded one iteration ago) and its first SA entry
    const bool have1 = i + stride < nq;
    uint64_t pred1 = narrow_finish(ix, x1, m1, pol.model);
    if (have1) pred1 = clamp_prediction(ix, pred1);
    else pred1 = pred0;
    const uint32_t idx1 = ld_u32_pol(ix.sa + pred1, pol.sa);
    // stage D: replay query t
    KmerQuery q;
    q.q = x0 << lsh;
    q.k = (uint32_t)ix.k;
    SaDirect sad;
    const long long r = pl_query_from<false, true>(ix, q, pred0, idx0, pol, sad);
    __stcs(out + i, r);
    x0 = x1; x1 = x2; x2 = x3;
    pred0 = pred1; idx0 = idx1; m1 = m2;
  }
}

// Result store shared by the production kernels.  slot != nullptr: the batch was partitioned (partition.cu); the
// query's position inside its chunk rides in the top 16 bits of the result word (an answer is -1 or < 2^32, so 48
// bits hold it) and the un-permute pass puts it back in the caller's order.
__device__ __forceinline__ void store_result(const IndexView& ix, long long* __restrict__ out,
                                             const uint16_t* __restrict__ slot, size_t i, long long r) {
  if (slot) r = (long long)(((unsigned long long)__ldcs(slot + i) << 48) | ((unsigned long long)r & 0xFFFFFFFFFFFFull));
  if (slot || (ix.hints & HINT_IO_STREAM)) __stcs(out + i, r);
  else out[i] = r;
}

// Sector-cached variant (default): the 8 suffix-array ranks around rev[predicted] come in one 256-bit load and stay
// in registers (query.cuh SaSector); works with either model layout.
template <int kMinBlocks, bool kLean>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_sector_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out,
                         const uint16_t* __restrict__ slot, unsigned long long* __restrict__ tiles) {
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  QueryCursor cur;
  for (cur.init(tiles); cur.more(nq); cur.next()) {
    const size_t i = cur.i;
    if (i >= nq) continue;  // ragged last tile
    const uint64_t x = (ix.hints & HINT_IO_STREAM) ? __ldcs(kmers + i) : __ldg(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
    long long r;
    if constexpr (kLean) {
      SaSector32 sa;
      sa.fill(ix, (uint32_t)pred, pol.sa);
      r = kmer_replay32<0, true>(ix, q.q, (uint32_t)pred, pol, sa);
    } else {
      SaSector sa;
      sa.fill(ix, pred, pol.sa);
      r = pl_query_from<false, false>(ix, q, pred, 0, pol, sa);
    }
    store_result(ix, out, slot, i, r);
  }
}

// Inline-prefix variant: every probe is one 16-byte ExtEntry {position, leading bases}; the packed genome is not
// touched at all (k <= ix.ext_bases).  For indexes whose genome does not fit L2 this halves the DRAM lines per query.
template <int kMinBlocks, bool kLean>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_inline_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out,
                         const uint16_t* __restrict__ slot, unsigned long long* __restrict__ tiles) {
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  QueryCursor cur;
  for (cur.init(tiles); cur.more(nq); cur.next()) {
    const size_t i = cur.i;
    if (i >= nq) continue;  // ragged last tile
    const uint64_t x = (ix.hints & HINT_IO_STREAM) ? __ldcs(kmers + i) : __ldg(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
    long long r;
    if constexpr (kLean) {
      SaNone32 none;
      r = kmer_replay32<1, true>(ix, q.q, (uint32_t)pred, pol, none);
    } else {
      SaDirect sad;
      r = pl_query_from<false, false, KmerQuery, SaDirect, true, 1>(ix, q, pred, 0, pol, sad);
    }
    store_result(ix, out, slot, i, r);
  }
}

// Rank-line variant: every probe is answered by a 32-byte sector {4 positions, 4 prefixes}, and the sectors a typical
// query needs share one 128-byte DRAM line (query.cuh SaPacked); the packed genome is read only for escaped entries
// and for queries longer than the entries' prefix.
template <int kMinBlocks, int kVariant>  // 0: general Replay, 1: lean replay, 2: lean replay + anchor line in shared memory,
                                         // 3: flat replay (tiling lines only)
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_packed_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out,
                         const uint16_t* __restrict__ slot, unsigned long long* __restrict__ tiles) {
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  QueryCursor cur;
  for (cur.init(tiles); cur.more(nq); cur.next()) {
    const size_t i = cur.i;
    if (i >= nq) continue;  // ragged last tile
    const uint64_t x = (ix.hints & HINT_IO_STREAM) ? __ldcs(kmers + i) : __ldg(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
    long long r;
    if constexpr (kVariant == 3) {
      r = kmer_replay_flat<true>(ix, q.q, (uint32_t)pred, pol);
    } else if constexpr (kVariant == 2) {
      __shared__ uint4 lines[kQueryThreads * kLineSlotU4];
      SaLine32 sa;
      sa.sm = lines + threadIdx.x * kLineSlotU4;
      sa.anchor(ix, (uint32_t)pred, pol.sa);
      r = kmer_replay32<2, true>(ix, q.q, (uint32_t)pred, pol, sa);
    } else if constexpr (kVariant == 1) {
      SaPacked32 sa;
      sa.anchor(ix, (uint32_t)pred);
      r = kmer_replay32<2, true>(ix, q.q, (uint32_t)pred, pol, sa);
    } else {
      SaPacked sa;
      sa.anchor(ix, pred);
      r = pl_query_from<false, false, KmerQuery, SaPacked, true, 2>(ix, q, pred, 0, pol, sa);
    }
    store_result(ix, out, slot, i, r);
  }
}

// In-order, software-pipelined kernel of the partitioned batch path (partition.cu).  Same replay as the kernels above;
// what differs is what is in flight while a lane replays query t of its warp's tile sequence:
//   * the warp's tiles come from the global in-order counter (QueryCursor explains why the order matters), claimed three
//     ahead;
//   * the k-mers of tile t+2 and the model checkpoints of tile t+1 are already requested, so the two dependent round trips
//     that head every query in the kernels above (k-mer -> checkpoints -> first suffix-array read; 14 % of all stall
//     samples in profiles/r1y) overlap with the probes of the previous tile;
//   * the slot of the query (partition.cu) is requested before the replay and consumed after it.
// Needs the narrow model layout.  kMode as in Replay: 0 = {suffix array sector, packed genome}, 1 = inline-prefix
// entries, 2 = rank lines, 3 = rank lines with the anchor line staged in shared memory (SaLine32; lean replay only),
// 4 = tiling rank lines answered by kmer_replay_flat (lean only), 5 = rank lines, lean replay, unfinished queries parked
// after three probes and resumed 32 at a time (needs the slot inside the k-mer word).
template <int kMinBlocks, int kMode, bool kLean>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_ordered_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out,
                          const uint16_t* __restrict__ slot, unsigned long long* __restrict__ tiles) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  // a partitioned batch has fewer than 2^32 queries (launch_partitioned_query): indices are 32-bit
  const uint32_t nq32 = (uint32_t)nq, last = nq32 - 1u;
  // tiles are claimed from the global in-order counter four at a time (128 consecutive queries per atomic): the claim
  // with its shuffles and reconvergence was 42 warp instructions per tile (ncu r2f, SASS page)
  constexpr unsigned kSpan = 4;
  uint32_t span_next = 0;
  unsigned span_left = 0;  // warp-uniform
  auto claim = [&]() -> uint32_t {
    if (span_left == 0) {
      unsigned long long a = 0;
      if (lane == 0) a = atomicAdd(tiles, 32ull * kSpan);
      a = __shfl_sync(0xffffffffu, a, 0);
      span_next = a < 0xFFFFFE00ull ? (uint32_t)a : 0xFFFFFE00u;  // past the end either way; keeps t + lane from wrapping
      span_left = kSpan;
    }
    const uint32_t t = span_next;
    span_next += 32u;
    span_left--;
    return t;
  };
  auto kmer_at = [&](uint32_t t) {  // past the end: the last k-mer again (loaded, predicted for, never answered)
    const uint32_t i = t + lane;
    return __ldcs(kmers + (i < last ? i : last));
  };
  // partition.cuh: the slot may ride in bits 50-63 of the k-mer word (uniform for the launch)
  const bool in_kmer = slot == slot_in_kmer_tag();
  const uint64_t kmask = in_kmer ? kSlotKmerMask : ~0ull;
  // kMode 5: per-warp queue of parked queries (structure of arrays: lane j of a drain reads word j of each array)
  __shared__ uint32_t park_s[(kLean && kMode == 5) ? kQueryThreads / 32 : 1][7][(kLean && kMode == 5) ? 64 : 1];
  uint32_t (*pk)[(kLean && kMode == 5) ? 64 : 1] = park_s[(kLean && kMode == 5) ? (threadIdx.x >> 5) : 0];
  unsigned parked = 0;  // warp-uniform
  Lean32 ps;
  bool pending = false;
  auto drain = [&](unsigned e) {  // resume parked query e in binarySearch and store its answer
    const uint64_t w = ((uint64_t)pk[1][e] << 32) | pk[0][e];
    Lean32 d;
    d.lo = pk[2][e];
    d.hi = pk[3][e];
    d.r = pk[4][e];
    const uint32_t c = pk[5][e];
    d.loLcp = c & 0xFFu;
    d.hiLcp = (c >> 8) & 0xFFu;
    d.start = (c >> 16) & 0xFFu;
    d.state = (int)(c >> 24);
    const uint32_t qi = pk[6][e];
    SaPacked32 sa;
    sa.abase = 0xFFFFFFF0u;  // no anchor line: every rank resolves to its own line
    sa.cur = 0xFFFFFFFFu;
    const long long r = kmer_replay32_tail<2, true>(ix, (w & kmask) << lsh, pol, sa, d);
    __stcs(out + qi, (long long)(((unsigned long long)(w >> kSlotShift) << 48) | ((unsigned long long)r & 0xFFFFFFFFFFFFull)));
  };
  uint32_t t0 = claim(), t1 = claim(), t2 = claim();
  if (t0 >= nq32) return;
  uint64_t x0 = kmer_at(t0), x1 = kmer_at(t1);
  NarrowPair m0 = narrow_load(ix, x0 & kmask, pol.model);
  while (t0 < nq32) {
    const uint64_t x2 = kmer_at(t2);
    const NarrowPair m1 = narrow_load(ix, x1 & kmask, pol.model);
    const uint32_t i = t0 + lane;
    pending = false;
    if (i < nq32) {
      const unsigned long long sl = in_kmer ? (unsigned long long)(x0 >> kSlotShift) : (unsigned long long)__ldcs(slot + i);
      const uint64_t pred = clamp_prediction(ix, narrow_finish(ix, x0 & kmask, m0, pol.model));
      KmerQuery q;
      q.q = (x0 & kmask) << lsh;
      q.k = (uint32_t)ix.k;
      long long r;
      if constexpr (kLean && kMode == 5) {
        SaPacked32 sa;
        sa.anchor(ix, (uint32_t)pred);
        pending = !kmer_replay32_head<2, true>(ix, q.q, (uint32_t)pred, pol, sa, ps, &r);
      } else if constexpr (kLean && kMode == 4) {
        r = kmer_replay_flat<true>(ix, q.q, (uint32_t)pred, pol);
      } else if constexpr (kLean && kMode == 3) {
        __shared__ uint4 lines[kQueryThreads * kLineSlotU4];
        SaLine32 sa;
        sa.sm = lines + threadIdx.x * kLineSlotU4;
        sa.anchor(ix, (uint32_t)pred, pol.sa);
        r = kmer_replay32<2, true>(ix, q.q, (uint32_t)pred, pol, sa);
      } else if constexpr (kLean && kMode == 2) {
        SaPacked32 sa;
        sa.anchor(ix, (uint32_t)pred);
        r = kmer_replay32<2, true>(ix, q.q, (uint32_t)pred, pol, sa);
      } else if constexpr (kLean && kMode == 1) {
        SaNone32 none;
        r = kmer_replay32<1, true>(ix, q.q, (uint32_t)pred, pol, none);
      } else if constexpr (kLean) {
        SaSector32 sa;
        sa.fill(ix, (uint32_t)pred, pol.sa);
        r = kmer_replay32<0, true>(ix, q.q, (uint32_t)pred, pol, sa);
      } else if constexpr (kMode >= 2) {
        SaPacked sa;
        sa.anchor(ix, pred);
        r = pl_query_from<false, false, KmerQuery, SaPacked, true, 2>(ix, q, pred, 0, pol, sa);
      } else if constexpr (kMode == 1) {
        SaDirect sad;
        r = pl_query_from<false, false, KmerQuery, SaDirect, true, 1>(ix, q, pred, 0, pol, sad);
      } else {
        SaSector sa;
        sa.fill(ix, pred, pol.sa);
        r = pl_query_from<false, false>(ix, q, pred, 0, pol, sa);
      }
      if (!(kLean && kMode == 5) || !pending)
        __stcs(out + i, (long long)((sl << 48) | ((unsigned long long)r & 0xFFFFFFFFFFFFull)));
    }
    if constexpr (kLean && kMode == 5) {
      // park what is still searching: after three probes about a third of the lanes are, and the binarySearch loop used to
      // run with ~10 of 32 lanes active (ncu r2k, SASS page: 375 of 726 warp instructions per tile at 10.9 lanes)
      const unsigned pmask = __ballot_sync(0xffffffffu, pending);
      if (pending) {
        const unsigned pos = parked + (unsigned)__popc(pmask & ((1u << lane) - 1u));
        pk[0][pos] = (uint32_t)x0;
        pk[1][pos] = (uint32_t)(x0 >> 32);
        pk[2][pos] = ps.lo;
        pk[3][pos] = ps.hi;
        pk[4][pos] = ps.r;
        pk[5][pos] = ps.loLcp | (ps.hiLcp << 8) | (ps.start << 16) | ((uint32_t)ps.state << 24);
        pk[6][pos] = t0 + lane;
      }
      parked += (unsigned)__popc(pmask);
      __syncwarp();
      if (parked >= 32u) {
        parked -= 32u;
        drain(parked + lane);
        __syncwarp();
      }
    }
    x0 = x1;
    x1 = x2;
    m0 = m1;
    t0 = t1;
    t1 = t2;
    t2 = claim();
  }
  if constexpr (kLean && kMode == 5) {
    if (lane < parked) drain(lane);
  }
}

// Lane-refill variant of the rank-line kernel.  In the kernels above a warp holds 32 queries and runs until the
// slowest of them is answered: the replay takes 1 to ~6 probes (mean 2.3), so on average barely half the lanes of an
// issued instruction do useful work (ncu: 17 active threads per warp instruction) and every query pays the dependent
// latency of the k-mer load, the model load and the longest probe chain in its warp.  Here a lane that finishes takes
// the next query at once:
//   * each warp owns a contiguous slice of the batch and reads it in 32-k-mer blocks, one coalesced load per block,
//     two blocks ahead (registers; lanes fetch their k-mer with a shuffle);
//   * every lane keeps a STANDBY query whose model checkpoints were requested when the standby slot was filled and are
//     consumed only when the current query finishes, at least one probe later -- so the k-mer and model loads never sit
//     on the critical path;
//   * one loop iteration = one probe (one rank-line sector) for every lane of the warp.
// Needs the narrow model layout.  Results are written straight to out[query index].
template <int kMinBlocks>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_packed_refill_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq,
                                long long* __restrict__ out) {
  using Rp = Replay<false, false, KmerQuery, SaPacked, true, 2>;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // slice of this warp: whole 32-query blocks, the last warp takes the ragged tail
  const size_t blocks_total = (nq + 31) >> 5;
  const size_t per = (blocks_total + warps - 1) / warps;
  const size_t begin = gw * per * 32 < nq ? gw * per * 32 : nq;
  const size_t end = (gw + 1) * per * 32 < nq ? (gw + 1) * per * 32 : nq;
  if (begin >= end) return;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  auto load_block = [&](size_t first) -> uint64_t {  // k-mer (first + lane), clamped inside the slice
    const size_t i = first + lane;
    return __ldcs(kmers + (i < end ? i : end - 1));
  };
  uint64_t blk0 = load_block(begin), blk1 = load_block(begin + 32), blk2 = load_block(begin + 64);
  size_t blk_first = begin;  // query index of blk0's lane 0
  size_t next = begin;       // next query index to hand out (warp-uniform)

  bool cur_valid = false, sb_valid = false;
  size_t cur_qi = 0, sb_qi = 0;
  uint64_t sb_x = 0;
  NarrowPair sb_m;
  sb_m.e0 = make_uint2(0, 0);
  sb_m.e1 = make_uint2(0, 0);
  KmerQuery q;
  q.q = 0;
  q.k = (uint32_t)ix.k;
  Rp rp;
  rp.begin(0);
  SaPacked sa;
  sa.abase = 0;
  sa.cur = ~0ull;

  for (;;) {
    // A: a lane without a current query promotes its standby (its checkpoints were requested an iteration ago or more)
    if (!cur_valid && sb_valid) {
      const uint64_t pred = clamp_prediction(ix, narrow_finish(ix, sb_x, sb_m, pol.model));
      q.q = sb_x << lsh;
      cur_qi = sb_qi;
      rp.begin(pred);
      sa.anchor(ix, pred);
      cur_valid = true;
      sb_valid = false;
    }
    // B: lanes with an empty standby slot take the next query indices, in lane order
    const unsigned want = __ballot_sync(0xffffffffu, !sb_valid);
    if (want && next < end) {
      const size_t mine = next + (size_t)__popc(want & lt_mask);
      const unsigned off = (unsigned)(mine - blk_first);  // < 64: at most 32 indices handed out per iteration
      const uint64_t a = __shfl_sync(0xffffffffu, blk0, (int)(off & 31u));
      const uint64_t b = __shfl_sync(0xffffffffu, blk1, (int)(off & 31u));
      if (!sb_valid && mine < end) {
        sb_x = off < 32u ? a : b;
        sb_qi = mine;
        sb_m = narrow_load(ix, sb_x, pol.model);  // asynchronous: not consumed before step A of a later iteration
        sb_valid = true;
      }
      next += (size_t)__popc(want);
      if (next > end) next = end;
      if (next - blk_first >= 32) {  // blk0 is used up: rotate, fetch two blocks ahead
        blk0 = blk1;
        blk1 = blk2;
        blk_first += 32;
        blk2 = load_block(blk_first + 64);
      }
    }
    // C: done when no lane holds a query any more
    if (!__any_sync(0xffffffffu, cur_valid || sb_valid)) break;
    // D: one probe
    if (cur_valid) {
      long long res;
      if (rp.step(ix, q, 0, pol, sa, &res)) {
        __stcs(out + cur_qi, res);
        cur_valid = false;
      }
    }
  }
}

// Traffic attribution (tools/ncu_hints.sh, SAPLING_B200_STAGES=1|2): the same kernel cut short after the model
// lookup (1) or after the suffix-array sector fetch (2); never used to answer queries.
__global__ void __launch_bounds__(kQueryThreads, 4)
kmer_query_stages_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out,
                         int stages) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const L2Policies pol = make_policies(ix.hints);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t x = __ldg(kmers + i);
    const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
    long long r = (long long)pred;
    if (stages >= 2) {
      SaSector sa;
      sa.fill(ix, pred, pol.sa);
      r = (long long)sa.ld(ix, pred, pol.sa);
    }
    out[i] = r;
  }
}

// Line-cached variant: the aligned 64-byte suffix-array line around rev[predicted] is fetched once per query
// into shared memory (query.cuh SaLine); works with either model layout.
template <int kMinBlocks>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_line_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, long long* __restrict__ out) {
  __shared__ uint4 lines[4 * kQueryThreads];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  SaLine<kQueryThreads> sa;
  sa.buf = lines + threadIdx.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t x = __ldcs(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
    sa.issue(ix, pred, pol.sa);
    sa.template wait<0>();
    __stcs(out + i, pl_query_from<false, false>(ix, q, pred, 0, pol, sa));
  }
}

// Line-cached + software-pipelined (narrow model layout only).  While a thread replays query t out of line
// buffer t&1, the cp.async of query t+1's suffix-array line into the other buffer, the model checkpoints of
// query t+2 and the k-mer of query t+3 are in flight.
template <int kMinBlocks>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_line_pipelined_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq,
                                 long long* __restrict__ out) {
  __shared__ uint4 lines[2][4 * kQueryThreads];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= nq) return;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  const size_t last = nq - 1;
  auto kmer_at = [&](size_t i) { return __ldcs(kmers + (i < last ? i : last)); };

  SaLine<kQueryThreads> sa0, sa1;
  sa0.buf = lines[0] + threadIdx.x;
  sa1.buf = lines[1] + threadIdx.x;
  uint64_t x0 = kmer_at(i0);
  uint64_t x1 = kmer_at(i0 + stride);
  uint64_t x2 = kmer_at(i0 + 2 * stride);
  uint64_t pred0 = clamp_prediction(ix, narrow_finish(ix, x0, narrow_load(ix, x0, pol.model), pol.model));
  sa0.issue(ix, pred0, pol.sa);
  NarrowPair m1 = narrow_load(ix, x1, pol.model);

  for (size_t i = i0; i < nq; i += stride) {
    const uint64_t x3 = kmer_at(i + 3 * stride);                 // k-mer of query t+3
    const NarrowPair m2 = narrow_load(ix, x2, pol.model);        // checkpoints of query t+2
    uint64_t pred1 = narrow_finish(ix, x1, m1, pol.model);       // prediction of query t+1 ...
    if (i + stride < nq) pred1 = clamp_prediction(ix, pred1);
    else pred1 = pred0;
    sa1.issue(ix, pred1, pol.sa);                                // ... and its suffix-array line
    sa0.template wait<1>();                                      // line of query t has landed
    KmerQuery q;
    q.q = x0 << lsh;
    q.k = (uint32_t)ix.k;
    __stcs(out + i, pl_query_from<false, false>(ix, q, pred0, 0, pol, sa0));
    x0 = x1; x1 = x2; x2 = x3;
    pred0 = pred1; m1 = m2;
    uint4* t = sa0.buf; sa0.buf = sa1.buf; sa1.buf = t;
    sa0.base = sa1.base;
  }
  sa0.template wait<0>();
}

__global__ void __launch_bounds__(kQueryThreads)
string_query_kernel(const IndexView ix, const uint64_t* __restrict__ words, const uint64_t* __restrict__ word_off,
                    const uint32_t* __restrict__ slens, const uint32_t* __restrict__ lengths,
                    const long long* __restrict__ kmers, size_t nq, long long* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    StringQuery q;
    q.w = words + word_off[i];
    q.slen_ = slens[i];
    q.length_ = lengths ? lengths[i] : slens[i];
    out[i] = pl_query<true>(ix, q, (uint64_t)kmers[i]);
  }
}

// The seed lookups of align.cpp seed_extend (:259-300) for a block of reads: one thread per (read, strand, seed).
// Seed position per :271-275, reverse complement per :241-256, kmerize + plQuery(query, val, k) :278-279, verification
// against the genome :283-285, then sa_pos = inverse-SA[ref_pos] and countHitsLeft/Right (:287-289, sapling_api.h:254-289)
// over the lcp>=k flags.  A seed holding a byte other than A/C/G/T can never verify in the reference (the compare at
// :284 is on raw bytes against a pure-ACGT genome), so it is answered -1 without a query.
__global__ void __launch_bounds__(kQueryThreads)
seed_kernel(const IndexView ix, const uint32_t* __restrict__ isa, const uint8_t* __restrict__ kflag,
            const char* __restrict__ reads, const uint64_t* __restrict__ off, size_t n_reads, uint32_t num_seeds,
            uint32_t maxHits, long long* __restrict__ ref_pos, uint32_t* __restrict__ sa_pos,
            uint32_t* __restrict__ left, uint32_t* __restrict__ right) {
  const size_t total = n_reads * 2 * (size_t)num_seeds;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint32_t k = (uint32_t)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const size_t r = t / (2 * (size_t)num_seeds);
    const uint32_t rem = (uint32_t)(t - r * 2 * (size_t)num_seeds);
    const uint32_t strand = rem / num_seeds, i = rem - strand * num_seeds;
    const uint64_t o = off[r], len = off[r + 1] - o;
    long long hit = -1;
    uint32_t rank = 0, lf = 0, rt = 0;
    if (len >= k) {
      const uint64_t last = len - k;
      uint64_t cur = 0;
      if (i == num_seeds - 1) cur = last;
      else if (i > 0) cur = last / (num_seeds - 1) * i;
      uint64_t x = 0;
      bool valid = true;
      for (uint32_t j = 0; j < k; j++) {
        const char c = strand ? reads[o + (len - 1 - (cur + j))] : reads[o + cur + j];
        uint32_t v;
        if (c == 'A') v = 0; else if (c == 'C') v = 1; else if (c == 'G') v = 2; else if (c == 'T') v = 3;
        else { v = 0; valid = false; }
        x = (x << 2) | (uint64_t)(strand ? 3u - v : v);  // complement: A<->T, C<->G
      }
      if (valid) {
        KmerQuery q;
        q.q = x << (64u - 2u * k);
        q.k = k;
        const uint64_t pred = clamp_prediction(ix, predict_rank(ix, x, pol.model));
        SaSector sa;
        sa.fill(ix, pred, pol.sa);
        const long long a = pl_query_from<false, false>(ix, q, pred, 0, pol, sa);
        if (a >= 0 && (uint64_t)a + k <= ix.n &&
            (load_bases_upto(ix.genome, (uint64_t)a, k) >> (64u - 2u * k)) == x) {
          hit = a;
          rank = isa[a];
          const uint64_t p = rank;
          uint32_t c = 0;
          for (; c < maxHits; c++)  // countHitsRight, sapling_api.h:254-263
            if ((uint64_t)c + p > ix.n - k || !kflag[c + p]) break;
          rt = c;
          for (c = 0; c < maxHits; c++)  // countHitsLeft, :283-289
            if (p < c || !kflag[p - c]) break;
          lf = c;
        }
      }
    }
    ref_pos[t] = hit;
    sa_pos[t] = rank;
    left[t] = lf;
    right[t] = rt;
  }
}

__global__ void predict_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq,
                               uint64_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride)
    out[i] = predict_rank(ix, kmers[i], make_policies(0).model);
}

// pos_j = splitmix64(seed + j) mod (n-k); optional 1-2 substitutions on odd j
__global__ void sample_kernel(const IndexView ix, uint64_t seed, uint64_t mut_seed, uint64_t first, size_t nq,
                              uint64_t* __restrict__ kmers) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned k = (unsigned)ix.k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t j = first + i;
    const uint64_t pos = splitmix64(seed + j) % (ix.n - k);
    uint64_t x = load_bases_upto(ix.genome, pos, k) >> (64u - 2u * k);
    if (mut_seed && (j & 1ull)) {
      const uint64_t h = splitmix64(mut_seed + j);
      const unsigned nsub = 1u + (unsigned)(h & 1ull);
      for (unsigned r = 0; r < nsub; r++) {
        const uint64_t hi = splitmix64(h + r + 1);
        const unsigned p = (unsigned)(hi % k);
        const unsigned sh = 2u * (k - 1u - p);
        const uint64_t old = (x >> sh) & 3ull;
        const uint64_t nw = (old + 1ull + (hi >> 32) % 3ull) & 3ull;
        x = (x & ~(3ull << sh)) | (nw << sh);
      }
    }
    kmers[i] = x;
  }
}

// the self-check of sapling_example.cpp:144-154 on the device
__global__ void verify_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, const long long* __restrict__ out,
                              size_t nq, unsigned long long* __restrict__ counters) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned k = (unsigned)ix.k;
  unsigned long long ok = 0, m1 = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const long long a = out[i];
    if (a == -1) { m1++; continue; }
    if ((uint64_t)a + k <= ix.n) {
      const uint64_t g = load_bases_upto(ix.genome, (uint64_t)a, k) >> (64u - 2u * k);
      if (g == kmers[i]) ok++;
    }
  }
  for (int o = 16; o; o >>= 1) {
    ok += __shfl_xor_sync(0xffffffffu, ok, o);
    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ok) atomicAdd(counters + 0, ok);
    if (m1) atomicAdd(counters + 1, m1);
  }
}

// P of SURVEY 8d, counted on the device: the number of getLcp calls (sapling_api.h:115-120) the REFERENCE makes for each
// query, i.e. the literal replay without the long-window shortcut; the final unverified rev[lo + 1] (:136,:247) is a
// suffix-array read, not a getLcp call.  bench.py uses the mean as `probes_per_query` on every rank.
__global__ void __launch_bounds__(kQueryThreads)
probe_count_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, unsigned long long* __restrict__ total) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const unsigned lsh = 64u - 2u * (unsigned)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  unsigned long long probes = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t x = __ldg(kmers + i);
    KmerQuery q;
    q.q = x << lsh;
    q.k = (uint32_t)ix.k;
    uint64_t pred = predict_rank(ix, x, pol.model);
    if (pred >= ix.n) pred = ix.n - 1;  // SURVEY H9, without bumping the out-of-range counter
    Replay<false, false, KmerQuery, SaDirect, false, 0> rp;
    rp.begin(pred);
    SaDirect sa;
    long long result;
    for (;;) {
      const bool final_read = rp.state == ST_FINAL;
      const bool done = rp.step(ix, q, 0, pol, sa, &result);
      if (!final_read) probes++;
      if (done) break;
    }
  }
  for (int o = 16; o; o >>= 1) probes += __shfl_xor_sync(0xffffffffu, probes, o);
  if ((threadIdx.x & 31) == 0 && probes) atomicAdd(total, probes);
}

// random 32-byte-sector gather: each thread chases nothing, it just issues independent sector
// reads at hashed addresses -- the "HBM random-sector roofline" denominator
__global__ void __launch_bounds__(256)
gather_kernel(const uint4* __restrict__ buf, uint64_t nsectors, uint64_t nloads, uint64_t salt,
              unsigned long long* __restrict__ sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloads; i += stride) {
    const uint64_t s = __umul64hi(splitmix64(salt + i), nsectors);  // uniform in [0, nsectors)
    const uint4 v = __ldg(buf + 2 * s);  // first 16 bytes of the sector: one sector transaction
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// generalised: each access reads `gran` contiguous bytes (32/64/128) at a random gran-aligned address;
// `chain` > 1 makes each thread follow a dependent chain of that many accesses (address of the next
// access derived from the loaded value), the access pattern of one query
template <int kGran>
__global__ void __launch_bounds__(256)
gather2_kernel(const uint4* __restrict__ buf, uint64_t nunits, uint64_t nthreads_work, int chain, uint64_t salt,
               unsigned long long* __restrict__ sink, unsigned nslices) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  // nslices > 1: block b only touches slice (b mod nslices) of the buffer (TLB-friendly, still DRAM-random)
  const uint64_t per_slice = nunits / nslices;
  const uint64_t slice_base = (uint64_t)(blockIdx.x % nslices) * per_slice;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nthreads_work; i += stride) {
    uint64_t h = splitmix64(salt + i);
    for (int c = 0; c < chain; c++) {
      const uint64_t u = slice_base + __umul64hi(h, per_slice);  // uniform in the slice
      const uint4* p = buf + u * (kGran / 16);
      unsigned v = 0;
#pragma unroll
      for (int j = 0; j < kGran / 16; j++) {
        const uint4 t = __ldg(p + j);
        v ^= t.x ^ t.y ^ t.z ^ t.w;
      }
      acc ^= v;
      h = splitmix64(h ^ v);  // next address depends on the data
    }
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

inline int query_grid(size_t nq, int blocks_per_sm) {
  size_t g = (nq + kQueryThreads - 1) / kQueryThreads;
  const size_t cap = (size_t)148 * blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// Experiment knob (tools/gpu_experiments.py): SAPLING_B200_QV = resident blocks per SM the kernel is
// compiled for (4: <=64 regs, 5, 6: <=40 regs, 8: <=32 regs).  Default chosen by measurement.
static int query_variant(const IndexView& ix, bool inline_layout, bool partitioned) {
  const char* e = getenv("SAPLING_B200_QV");  // read per launch so one process can sweep variants
  // Measured (profiles/r1_experiments.md): while the genome and model mostly hit L2 more resident warps help (5
  // blocks/SM at c2: +8 %); once every access is a DRAM line and a TLB miss fewer do better (3 blocks/SM at c3: +9 %).
  // The inline-prefix kernel is best at 4 (profiles/r1_c3_inline.md).
  // A partitioned batch (in-order tiles, the slice of the index in L2) is latency- and issue-bound rather than DRAM-bound:
  // 5 blocks/SM (gpurun r2d / r2e: c2 1.55 against 1.63 ms per 50 M, c3 10.1 against 10.7 ms per 250 M).
  int v = e ? atoi(e) : (partitioned ? 5 : inline_layout ? 4 : (ix.n > 1000000000ull ? 3 : 5));
  if (v != 2 && v != 3 && v != 4 && v != 5 && v != 6 && v != 8) v = 4;
  return v;
}

// name_out != nullptr: only report which kernel a batch would run on (introspection for bench.py), launch nothing
// d_slot != nullptr: partitioned batch (partition.cu) -- results carry their chunk slot in the top 16 bits
int launch_kmer_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, cudaStream_t st,
                      const char** name_out, const uint16_t* d_slot, unsigned long long* d_tiles) {
  if (nq == 0 && !name_out) return 0;
  const char* gm = getenv("SAPLING_B200_GRID_MULT");  // grid = 148 * blocks/SM * mult (experiment knob)
  const int mult = gm ? (atoi(gm) > 0 ? atoi(gm) : 2) : 2;
  // Measured on the c2 workload (profiles/r1_experiments.md): the plain one-thread-per-query kernel at 4 blocks/SM
  // and two waves is the fastest; the software-pipelined and line-cached variants stay selectable for experiments.
  const char* pe = getenv("SAPLING_B200_PIPELINE");  // 1 = software-pipelined
  const char* le = getenv("SAPLING_B200_LINE");      // 1 = suffix-array line cached in shared memory
  const char* se = getenv("SAPLING_B200_SECTOR");    // 0 = per-read suffix-array gathers (the round-1 baseline kernel)
  const bool pipelined = ix.narrow != nullptr && pe && atoi(pe) == 1;
  const bool line = le && atoi(le) == 1;
  const bool sector = !(se && atoi(se) == 0) && !pipelined && !line;
  const char* ie = getenv("SAPLING_B200_INLINE_QUERY");  // 0 = ignore the inline-prefix array even if resident
  const bool inl = ix.ext != nullptr && ix.k <= ix.ext_bases && !(ie && atoi(ie) == 0);
  const char* pe2 = getenv("SAPLING_B200_PACKED_QUERY");  // 0 = ignore the rank lines even if resident
  const bool packed = ix.packed != nullptr && !(pe2 && atoi(pe2) == 0);
  // 1 = lane-refill kernel.  Measured slower than one query per lane per pass at every size (profiles/r1v_layouts.md:
  // c2 13.8 vs 18.2 G q/s, c3 9.9 vs 11.3), so it is opt-in.
  const char* re = getenv("SAPLING_B200_REFILL");
  const bool refill = packed && ix.narrow != nullptr && re && atoi(re) == 1 && !d_slot;
  const int qv = query_variant(ix, inl || packed, d_slot != nullptr);
  const char* lne = getenv("SAPLING_B200_LEAN");  // 0 = the general Replay instead of kmer_replay32 (A/B measurements)
  const bool lean = lean_eligible(ix) && !(lne && atoi(lne) == 0);
  const char* lse = getenv("SAPLING_B200_LINE_SMEM");  // 1 = anchor line staged in shared memory (measured slower: opt-in)
  const bool line_smem = lean && packed && lse && atoi(lse) == 1;
  // 1 = kmer_replay_flat instead of kmer_replay32.  Measured slower (gpurun r2g: c2 1.72 against 1.55 ms per 50 M
  // partitioned queries, c3 10.7 against 10.0 ms per 250 M): a lane executes ~155 instructions per probe where the branchy
  // replay executes ~110 on its path, and that costs more than the convergence wins back.  Opt-in.
  const char* fe = getenv("SAPLING_B200_FLAT");
  const bool flat = lean && packed && ix.packed_shift == 4 && !line_smem && fe && atoi(fe) == 1;
  if (const char* sg = getenv("SAPLING_B200_STAGES")) {
    if (atoi(sg) == 1 || atoi(sg) == 2) {
      kmer_query_stages_kernel<<<query_grid(nq, 4 * mult), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, atoi(sg));
      SB_CUDA_CHECK(cudaGetLastError());
      return 0;
    }
  }
  const char* oe = getenv("SAPLING_B200_ORDERED_PIPE");  // 0 = in-order tiles without the software pipeline
  const bool ordered = d_tiles && d_slot && ix.narrow != nullptr && (packed || inl || sector) && !(oe && atoi(oe) == 0);
  if (name_out) {
    *name_out = ordered ? "kmer_query_ordered_kernel" : refill ? "kmer_query_packed_refill_kernel" : packed ? "kmer_query_packed_kernel" : inl ? "kmer_query_inline_kernel" : sector ? "kmer_query_sector_kernel"
                : (line && pipelined) ? "kmer_query_line_pipelined_kernel" : line ? "kmer_query_line_kernel"
                : pipelined ? "kmer_query_pipelined_kernel" : "kmer_query_kernel";
    return qv;
  }
#define SB_LAUNCH(kernel, bps) kernel<bps><<<query_grid(nq, bps * mult), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out)
#define SB_LAUNCH_S(kernel, bps)                                                                                       \
  do {                                                                                                                \
    if (lean)                                                                                                         \
      kernel<bps, true><<<query_grid(nq, bps * (d_tiles ? 1 : mult)), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out,  \
                                                                                              d_slot, d_tiles);      \
    else                                                                                                              \
      kernel<bps, false><<<query_grid(nq, bps * (d_tiles ? 1 : mult)), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, \
                                                                                               d_slot, d_tiles);     \
  } while (0)
  if (d_slot == slot_in_kmer_tag() && !ordered) {
    set_error("slots inside the k-mer words are only read by the in-order pipelined kernel");
    return -1;
  }
  if (d_slot && !(packed || inl || sector)) {
    set_error("partitioned batches need the sector, inline or rank-line kernel");
    return -1;
  }
  if (ordered) {
#define SB_LAUNCH_O(bps, mode)                                                                                       \
  do {                                                                                                               \
    if (lean)                                                                                                        \
      kmer_query_ordered_kernel<bps, mode, true><<<query_grid(nq, bps), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, \
                                                                                                d_slot, d_tiles);   \
    else                                                                                                             \
      kmer_query_ordered_kernel<bps, mode, false><<<query_grid(nq, bps), kQueryThreads, 0, st>>>(                    \
          ix, d_kmers, nq, d_out, d_slot, d_tiles);                                                                  \
  } while (0)
    // parked binarySearch tails: rank lines, lean replay, slots inside the k-mer words (k <= 25).  MEASURED SLOWER once the
    // binarySearch loop had been specialised (gpurun r2o: c2 1.26 against 1.13 ms per 50 M, c3 8.2 against 7.6 ms per 250 M:
    // the queue traffic, the lost sector in registers and the scattered result stores cost more than the ~10 idle lanes
    // of a ~60-instruction loop body).  Opt-in: SAPLING_B200_PARK=1.
    const char* pke = getenv("SAPLING_B200_PARK");
    const bool park = lean && packed && !flat && !line_smem && d_slot == slot_in_kmer_tag() && pke && atoi(pke) == 1;
    const int mode = packed ? (park ? 5 : flat ? 4 : line_smem ? 3 : 2) : inl ? 1 : 0;
    switch (qv * 10 + mode) {
      case 35: SB_LAUNCH_O(3, 5); break;
      case 45: SB_LAUNCH_O(4, 5); break;
      case 55: SB_LAUNCH_O(5, 5); break;
      case 65: SB_LAUNCH_O(6, 5); break;
      case 62: SB_LAUNCH_O(6, 2); break;
      case 34: SB_LAUNCH_O(3, 4); break;
      case 44: SB_LAUNCH_O(4, 4); break;
      case 54: SB_LAUNCH_O(5, 4); break;
      case 64: SB_LAUNCH_O(6, 4); break;
      case 33: SB_LAUNCH_O(3, 3); break;
      case 43: SB_LAUNCH_O(4, 3); break;
      case 53: SB_LAUNCH_O(5, 3); break;
      case 30: SB_LAUNCH_O(3, 0); break;
      case 31: SB_LAUNCH_O(3, 1); break;
      case 32: SB_LAUNCH_O(3, 2); break;
      case 50: SB_LAUNCH_O(5, 0); break;
      case 51: SB_LAUNCH_O(5, 1); break;
      case 52: SB_LAUNCH_O(5, 2); break;
      case 41: SB_LAUNCH_O(4, 1); break;
      case 42: SB_LAUNCH_O(4, 2); break;
      default: SB_LAUNCH_O(4, 0); break;
    }
#undef SB_LAUNCH_O
    SB_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  if (refill) {
    // persistent: every warp streams through one contiguous slice of the batch, so one resident wave (unless
    // SAPLING_B200_GRID_MULT says otherwise)
#define SB_LAUNCH_R(bps) \
  kmer_query_packed_refill_kernel<bps><<<query_grid(nq, bps * (gm ? mult : 1)), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out)
    switch (qv) {
      case 2: SB_LAUNCH_R(2); break;
      case 3: SB_LAUNCH_R(3); break;
      case 5: SB_LAUNCH_R(5); break;
      default: SB_LAUNCH_R(4); break;
    }
#undef SB_LAUNCH_R
  } else if (packed) {
#define SB_LAUNCH_P(bps)                                                                                             \
  do {                                                                                                               \
    const int g = query_grid(nq, bps * (d_tiles ? 1 : mult));                                                        \
    if (flat) kmer_query_packed_kernel<bps, 3><<<g, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);  \
    else if (line_smem) kmer_query_packed_kernel<bps, 2><<<g, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles); \
    else if (lean) kmer_query_packed_kernel<bps, 1><<<g, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles); \
    else kmer_query_packed_kernel<bps, 0><<<g, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);       \
  } while (0)
    switch (qv) {
      case 3: SB_LAUNCH_P(3); break;
      case 5: SB_LAUNCH_P(5); break;
      case 6: SB_LAUNCH_P(6); break;
      default: SB_LAUNCH_P(4); break;
    }
#undef SB_LAUNCH_P
  } else if (inl) {
    switch (qv) {
      case 3: SB_LAUNCH_S(kmer_query_inline_kernel, 3); break;
      case 5: SB_LAUNCH_S(kmer_query_inline_kernel, 5); break;
      case 6: SB_LAUNCH_S(kmer_query_inline_kernel, 6); break;
      default: SB_LAUNCH_S(kmer_query_inline_kernel, 4); break;
    }
  } else if (sector) {
    switch (qv) {
      case 3: SB_LAUNCH_S(kmer_query_sector_kernel, 3); break;
      case 5: SB_LAUNCH_S(kmer_query_sector_kernel, 5); break;
      case 6: SB_LAUNCH_S(kmer_query_sector_kernel, 6); break;
      default: SB_LAUNCH_S(kmer_query_sector_kernel, 4); break;
    }
  } else if (line && pipelined) {
    switch (qv) {
      case 3: SB_LAUNCH(kmer_query_line_pipelined_kernel, 3); break;
      case 5: SB_LAUNCH(kmer_query_line_pipelined_kernel, 5); break;
      case 6: SB_LAUNCH(kmer_query_line_pipelined_kernel, 6); break;
      default: SB_LAUNCH(kmer_query_line_pipelined_kernel, 4); break;
    }
  } else if (line) {
    switch (qv) {
      case 3: SB_LAUNCH(kmer_query_line_kernel, 3); break;
      case 5: SB_LAUNCH(kmer_query_line_kernel, 5); break;
      case 6: SB_LAUNCH(kmer_query_line_kernel, 6); break;
      case 8: SB_LAUNCH(kmer_query_line_kernel, 8); break;
      default: SB_LAUNCH(kmer_query_line_kernel, 4); break;
    }
  } else if (pipelined) {
    switch (qv) {
      case 3: SB_LAUNCH(kmer_query_pipelined_kernel, 3); break;
      case 5: SB_LAUNCH(kmer_query_pipelined_kernel, 5); break;
      case 6: SB_LAUNCH(kmer_query_pipelined_kernel, 6); break;
      default: SB_LAUNCH(kmer_query_pipelined_kernel, 4); break;
    }
  } else {
    switch (qv) {
      case 5: SB_LAUNCH(kmer_query_kernel, 5); break;
      case 6: SB_LAUNCH(kmer_query_kernel, 6); break;
      case 8: SB_LAUNCH(kmer_query_kernel, 8); break;
      default: SB_LAUNCH(kmer_query_kernel, 4); break;
    }
  }
#undef SB_LAUNCH
#undef SB_LAUNCH_S
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_string_query(const IndexView& ix, const uint64_t* d_words, const uint64_t* d_word_off,
                        const uint32_t* d_slens, const uint32_t* d_lengths, const long long* d_kmers, size_t nq,
                        long long* d_out, cudaStream_t st) {
  if (nq == 0) return 0;
  string_query_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(ix, d_words, d_word_off, d_slens, d_lengths,
                                                                   d_kmers, nq, d_out);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_seeds(const IndexView& ix, const uint32_t* d_isa, const uint8_t* d_kflag, const char* d_reads,
                 const uint64_t* d_off, size_t n_reads, uint32_t num_seeds, uint32_t maxHits, long long* d_ref_pos,
                 uint32_t* d_sa_pos, uint32_t* d_left, uint32_t* d_right, cudaStream_t st) {
  const size_t total = n_reads * 2 * (size_t)num_seeds;
  if (total == 0) return 0;
  seed_kernel<<<query_grid(total, 8), kQueryThreads, 0, st>>>(ix, d_isa, d_kflag, d_reads, d_off, n_reads, num_seeds,
                                                              maxHits, d_ref_pos, d_sa_pos, d_left, d_right);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_predict(const IndexView& ix, const uint64_t* d_kmers, size_t nq, uint64_t* d_out, cudaStream_t st) {
  if (nq == 0) return 0;
  predict_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_sample(const IndexView& ix, uint64_t seed, uint64_t mut_seed, uint64_t first, size_t nq,
                  uint64_t* d_kmers, cudaStream_t st) {
  if (nq == 0) return 0;
  sample_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(ix, seed, mut_seed, first, nq, d_kmers);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_verify(const IndexView& ix, const uint64_t* d_kmers, const long long* d_out, size_t nq,
                  unsigned long long* d_counters, cudaStream_t st) {
  if (nq == 0) return 0;
  verify_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(ix, d_kmers, d_out, nq, d_counters);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_probe_count(const IndexView& ix, const uint64_t* d_kmers, size_t nq, unsigned long long* d_total,
                       cudaStream_t st) {
  if (nq == 0) return 0;
  probe_count_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_total);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int run_gather_bench(uint64_t bytes, uint64_t n_loads, int reps, double* gbps) {
  if (bytes < (1ull << 20)) bytes = 1ull << 20;
  void* buf = nullptr;
  unsigned long long* sink = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&buf, bytes));
  SB_CUDA_CHECK(cudaMalloc(&sink, 8));
  SB_CUDA_CHECK(cudaMemset(buf, 0x5A, bytes));
  SB_CUDA_CHECK(cudaMemset(sink, 0, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const uint64_t nsect = bytes / 32;
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
    gather_kernel<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(buf), nsect, n_loads, 0x1234ull + (uint64_t)r * n_loads, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(buf); cudaFree(sink); SB_CUDA_CHECK(e); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)n_loads * 32.0 / (ms * 1e-3) / 1e9;
    if (r > 0 && g > best) best = g;  // first rep is warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  if (gbps) *gbps = best;
  return 0;
}

int run_gather_bench2(uint64_t bytes, uint64_t n_access, int gran, int chain, int blocks_per_sm, int reps,
                      double* gacc_per_s) {
  if (bytes < (1ull << 20)) bytes = 1ull << 20;
  if (chain < 1) chain = 1;
  void* buf = nullptr;
  unsigned long long* sink = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&buf, bytes));
  SB_CUDA_CHECK(cudaMalloc(&sink, 8));
  SB_CUDA_CHECK(cudaMemset(buf, 0x5A, bytes));
  SB_CUDA_CHECK(cudaMemset(sink, 0, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const uint64_t nunits = bytes / (uint64_t)gran;
  const uint64_t work = n_access / (uint64_t)chain;
  const int grid = 148 * blocks_per_sm;
  const char* se = getenv("SAPLING_B200_GATHER_SLICES");  // experiment knob, see gather2_kernel
  const unsigned nslices = se && atoi(se) > 0 ? (unsigned)atoi(se) : 1u;
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    const uint64_t salt = 0x9999ull + (uint64_t)r * work;
    cudaEventRecord(e0);
    if (gran == 16) gather2_kernel<16><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink, nslices);
    else if (gran == 32) gather2_kernel<32><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink, nslices);
    else if (gran == 64) gather2_kernel<64><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink, nslices);
    else gather2_kernel<128><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink, nslices);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(buf); cudaFree(sink); SB_CUDA_CHECK(e); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)(work * (uint64_t)chain) / (ms * 1e-3) / 1e9;
    if (r > 0 && g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  if (gacc_per_s) *gacc_per_s = best;
  return 0;
}

}  // namespace sb

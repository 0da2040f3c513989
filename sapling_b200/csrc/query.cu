// query.cu -- query kernels and their launchers.
//
// Hot path: kmer_query_ordered_kernel (a partitioned batch, partition.cu) and kmer_query_kernel (a batch in the caller's
// order).  One lane per query; per query: queryPiecewiseLinear from the narrow model (common.cuh), then kmer.cuh -- the
// sector of the predicted rank classified against the query, its neighbours only when needed, the reference's control flow
// replayed in registers.  Everything a query reads is a read-only 32-byte sector gather (ld.global.nc.v8); DESIGN.md has
// the roofline, profiles/ the ncu evidence.
#include "build.cuh"
#include "common.cuh"
#include "kmer.cuh"
#include "query.cuh"

namespace sb {

namespace {

constexpr int kQueryThreads = 256;

// Result word of a partitioned batch: the query's slot inside its chunk (partition.cu) rides in the top 16 bits (an answer
// is -1 or < 2^32, so 48 bits hold it); the un-permute pass puts the answers back in the caller's order.
__device__ __forceinline__ long long slot_word(unsigned long long slot, long long r) {
  return (long long)((slot << 48) | ((unsigned long long)r & 0xFFFFFFFFFFFFull));
}

// A batch in the caller's order: grid-stride, k-mer reads and result writes coalesced.  Out = long long (the reference's
// return type) or uint32_t (0xFFFFFFFF for -1; the narrow download format of the host path).
template <int kMinBlocks, bool kTies, typename Out>
__global__ void __launch_bounds__(kQueryThreads, kMinBlocks)
kmer_query_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, Out* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const L2Policies pol = make_policies(ix.hints);
  const uint64_t kmask = ix.k >= 32 ? ~0ull : ((1ull << (2 * ix.k)) - 1ull);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const uint64_t x = __ldcs(kmers + i) & kmask;  // bits above 2k are not part of a k-mer: ignored, never indexed with
    const uint32_t pred = (uint32_t)clamp_prediction(ix, predict_rank(ix, x, pol.model));
    const long long r = answer_kmer_fast<kTies>(ix, x, pred, pol);
    __stcs(out + i, (Out)r);
  }
}

// The two checkpoints of a k-mer's bucket, in either model layout, requested one tile ahead of their use.
template <bool kNarrow>
struct ModelPair;
template <>
struct ModelPair<true> {
  NarrowPair p;
  __device__ __forceinline__ void load(const IndexView& ix, uint64_t x, uint64_t pol) { p = narrow_load(ix, x, pol); }
  __device__ __forceinline__ uint64_t predict(const IndexView& ix, uint64_t x, uint64_t pol) const {
    return narrow_finish(ix, x, p, pol);
  }
};
template <>
struct ModelPair<false> {
  longlong2 lo, hi;
  __device__ __forceinline__ void load(const IndexView& ix, uint64_t x, uint64_t pol) {
    const uint64_t b = x >> ix.shift;
    lo = ld_s64x2_pol(reinterpret_cast<const longlong2*>(ix.model + b), pol);
    hi = ld_s64x2_pol(reinterpret_cast<const longlong2*>(ix.model + b + 1), pol);
  }
  __device__ __forceinline__ uint64_t predict(const IndexView&, uint64_t x, uint64_t) const {
    return interpolate((long long)x, lo.x, lo.y, hi.x, hi.y);
  }
};

// In-order, software-pipelined kernel of the partitioned batch path (partition.cu).  The batch arrives bucketed by the top
// bits of the k-mer and must be WALKED IN ORDER for that to pay: warps claim tiles of 32 consecutive queries from a global
// counter (four tiles per atomic), so the ~150 k queries in flight on the GPU always fall into one or two slices of the
// index -- the slice of the rank lines and of the model is L2 resident, a DRAM line is filled once per batch.  While a lane
// answers its query of tile t, the k-mers of tile t+3, the model checkpoints of t+2 and the prediction and sector of t+1
// are on their way, so the dependent round trips that head every query (k-mer -> checkpoints -> first sector) overlap with
// the previous tiles.  The slot of a query rides in bits 50-63 of its k-mer word (k <= 25) or comes from a side array.
//
// Lane occupancy.  A third of the queries are answered by the sector of their predicted rank, nearly all the others by
// one neighbouring sector; 1 % of present k-mers (more of absent ones) need a third sector or more (errors beyond the 95 %
// bounds, match runs crossing a sector end, absent k-mers far from their prediction).  A warp that loops until its slowest
// lane is done spends most of its instructions with 2-4 lanes alive (ncu, gpurun s2: 13.7 of 32 lanes active per
// instruction).  So a tile gets exactly TWO classification rounds in place, decided by the two-sector shortcut (kmer.cuh:
// a handful of compares, no search state); what is still undecided is pushed onto the warp's own stack in shared memory,
// and whenever 32 have piled up the warp pops them and answers them with the general search, every lane busy at the start.
// No block-wide barrier: the warps stay independent.  (ncu r3e: 26 of 32 lanes active per instruction.)
constexpr int kWarpsPerBlock = kQueryThreads / 32;
constexpr int kStackCap = 64;  // at most 31 left over + 32 pushed by a tile
struct TailStacks {            // per warp: the queries the two-sector shortcut left undecided (k-mer word, prediction, index)
  uint32_t x_lo[kWarpsPerBlock][kStackCap], x_hi[kWarpsPerBlock][kStackCap];
  uint32_t pred[kWarpsPerBlock][kStackCap], idx[kWarpsPerBlock][kStackCap];
};

// The k-mer stream goes through shared memory: ptxas hoists the register rotation x2 = x3 to just behind the load of x3, so
// the warp waits out the DRAM round trip of a k-mer it does not need for another tile (ncu s10, SASS page: 15.5 % of all
// stall samples on that one MOV).  An asynchronous copy has no destination register to wait on: one coalesced LDGSTS per
// tile into a 2-stage ring of the warp, read back a tile later (gpurun s13: 5.78 -> 5.60 ms at c3, 1.086 -> 1.041 at c2).
// The per-lane gathers -- checkpoints, sectors -- stay register loads: as scattered LDGSTS they cost more in the
// shared-memory pipe than they save (gpurun s12: 8.4 ms).
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// kOut: what a finished query writes.  kOutSlotWord: a partitioned batch -- slot word at the query's partitioned index
// (slot_word above; the un-permute pass restores the caller's order).  kOutInt64 / kOutU32: a batch in the caller's order
// answered by the same pipeline (an index small enough to stay in L2 needs no partitioning, but the dependent reads of a
// query want the look-ahead just the same: 5 M queries at 10 Mbp take 0.36 ms through kmer_query_kernel) -- the answer
// itself at the query's index, -1 or 0xFFFFFFFF for "not found".
enum { kOutSlotWord = 0, kOutInt64 = 1, kOutU32 = 2 };
constexpr int kOrderedBlocks = 4;  // resident blocks per SM: 64 registers, no spills (3: 6.8 ms, 5: spills, 6.8 ms at c3)
template <bool kTies, bool kNarrow, int kOut>
__global__ void __launch_bounds__(kQueryThreads, kOrderedBlocks)
kmer_query_ordered_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq, void* __restrict__ out_raw,
                          const uint16_t* __restrict__ slot, unsigned long long* __restrict__ tiles) {
  __shared__ TailStacks stacks;
  __shared__ uint64_t kmer_ring[kWarpsPerBlock][2][32];  // the k-mers of tile t2 (landed) and of t3 (in flight)
  __shared__ uint4 near_ring[kWarpsPerBlock][2][32];  // second-round sectors of the current tile, in two halves
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const L2Policies pol = make_policies(ix.hints);
  // a partitioned batch has fewer than 2^32 queries (launch_partitioned_query): indices are 32-bit
  const uint32_t nq32 = (uint32_t)nq, last = nq32 - 1u;
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
  auto kmer_request = [&](uint32_t t, unsigned stage) {
    const uint32_t i = t + lane;
    cp_async8(&kmer_ring[warp][stage][lane], kmers + (i < last ? i : last));
    cp_async_commit();
  };
  const bool in_kmer = slot == slot_in_kmer_tag();
  // bits above 2k are not part of a k-mer (and bits 50-63 may carry the slot): never indexed with
  const uint64_t kmask = ix.k >= 32 ? ~0ull : ((1ull << (2 * ix.k)) - 1ull);
  auto store = [&](uint32_t i, uint64_t xw, long long r) {
    if (kOut == kOutInt64) {
      __stcs(static_cast<long long*>(out_raw) + i, r);
    } else if (kOut == kOutU32) {
      __stcs(static_cast<uint32_t*>(out_raw) + i, (uint32_t)r);
    } else {
      const unsigned long long sl = in_kmer ? (unsigned long long)(xw >> kSlotShift) : (unsigned long long)__ldcs(slot + i);
      __stcs(static_cast<long long*>(out_raw) + i, slot_word(sl, r));
    }
  };
  unsigned stacked = 0;  // warp-uniform: entries on this warp's stack
  auto push = [&](bool pending, uint64_t xw, uint32_t pred, uint32_t i) {
    const unsigned m = __ballot_sync(0xffffffffu, pending);
    if (pending) {
      const unsigned e = stacked + (unsigned)__popc(m & lt_mask);
      stacks.x_lo[warp][e] = (uint32_t)xw;
      stacks.x_hi[warp][e] = (uint32_t)(xw >> 32);
      stacks.pred[warp][e] = pred;
      stacks.idx[warp][e] = i;
    }
    stacked += (unsigned)__popc(m);
    __syncwarp();
  };
  // pop the top `m` (<= 32) entries and answer them with the general search
  auto drain = [&](unsigned m) {
    stacked -= m;
    if (lane < m) {
      const unsigned e = stacked + lane;
      const uint64_t xw = ((uint64_t)stacks.x_hi[warp][e] << 32) | stacks.x_lo[warp][e];
      const uint32_t i = stacks.idx[warp][e];
      store(i, xw, answer_kmer<kTies>(ix, xw & kmask, stacks.pred[warp][e], pol));
    }
    __syncwarp();
  };

  // Pipeline, in tiles ahead of the one being answered: k-mers 3, model checkpoints 2, prediction AND ITS SECTOR 1: the first
  // read of a query is the one that goes to DRAM (the later ones stay in the same 128-byte line), and it is in registers a
  // whole tile before the lane classifies it.  (Measured, gpurun s9, c3, ms per 250 M queries at 4 blocks per SM: nothing
  // ahead 6.17, an L2 prefetch of the sector 5.83, the sector itself 5.74.  The neighbour a second round is most likely to
  // ask for, prefetched into L1 along with it: 5.81 against 5.60, gpurun s14 -- one more sector request per query costs more
  // than the shorter wait of the 43 % of lanes that use it.)
  auto predict = [&](uint64_t xw, const ModelPair<kNarrow>& m, bool real, U32x8* sector) -> uint32_t {
    uint64_t p = m.predict(ix, xw & kmask, pol.model);
    if (real) p = clamp_prediction(ix, p);  // counts predictions past the last rank: real queries only
    else if (p >= ix.n) p = ix.n - 1;
    *sector = load_sector(ix, (uint32_t)(p >> 2), pol);
    return (uint32_t)p;
  };
  uint32_t t0 = claim(), t1 = claim(), t2 = claim(), t3 = claim();
  if (t0 >= nq32) return;
  uint64_t x0 = kmer_at(t0), x1 = kmer_at(t1), x2 = 0;
  unsigned kq = 0;  // ring stage of the k-mers of t2
  kmer_request(t2, kq);
  ModelPair<kNarrow> m1;
  uint32_t pred0;
  U32x8 sec0;
  {
    ModelPair<kNarrow> m0;
    m0.load(ix, x0 & kmask, pol.model);
    m1.load(ix, x1 & kmask, pol.model);
    pred0 = predict(x0, m0, t0 + lane < nq32, &sec0);
  }
  while (t0 < nq32) {
    uint32_t pred1 = 0;
    U32x8 sec1;
    ModelPair<kNarrow> m2;
    // what the next tiles need: the k-mers of t2 out of the ring and those of t3 asked for, checkpoints for t2, prediction
    // and sector for t1 -- done BETWEEN the request of a second-round sector and its use (below).  Not with ties (k longer
    // than the lines' prefixes: every match waits for a genome read inside round 1): there the look-ahead goes first, as
    // its requests then travel during that wait (k = 31, c4: 15.0 ms with the look-ahead first, 18.3 behind round 1).
    constexpr bool kLate = !kTies;
    auto look_ahead = [&]() {
      if (kLate) cp_async_wait_but_last();  // the k-mers of t2, asked for a tile ago (the last group: this tile's neighbours)
      else cp_async_wait_all();
      x2 = kmer_ring[warp][kq][lane];
      kq ^= 1u;
      kmer_request(t3, kq);
      m2.load(ix, x2 & kmask, pol.model);
      pred1 = predict(x1, m1, t1 + lane < nq32, &sec1);
    };
    if (!kLate) look_ahead();
    const uint32_t i = t0 + lane;
    // Every phase below is its own `if`: the lanes that need it meet there again whatever they did before.
    const bool active = i < nq32;
    bool done = false;
    int st = 2;  // 0: bounds final, 1: wants the neighbour, 2: left to the general search
    long long r = -1;
    uint32_t pred = 0, neighbour = 0;
    KmerKey key;
    key.q = key.qlo = key.qhi = 0;
    Bounds b;
    b.lb = b.ub = 0;
    Sector s0;
    s0.s = s0.c = s0.m = 0;
    if (active) {  // round 1: the sector of the predicted rank
      pred = pred0;
      key = make_key<kTies>(ix, x0 & kmask);
      uint32_t pos[4], idx = 0;
      s0 = classify_loaded<kTies>(ix, key, pred >> 2, sec0, pol, pos);
      done = direct_match(pred, s0, pos, &idx);  // :164
      r = (long long)idx;
      st = two_sector_first(ix, s0, &b, &neighbour);
    }
    if (kLate) {
      // round 2, asked for here and used after the look-ahead: the ~100 instructions of the next tiles' prologue (times
      // the other warps of the scheduler) run while the neighbour sector travels from L2.  Loaded in place it was the one
      // wait of the loop that nothing overlapped (ncu s13: 25 % of all stall samples on its first use; gpurun s16: 5.65 ->
      // 5.53 ms at c3, 5.72 -> 5.61 at c4, 1.04 -> 1.05 at c2).  Through shared memory: a register destination would be
      // live across the whole look-ahead.
      const bool second = active && !done && st == 1;
      if (second) {
        const uint32_t* p = ix.lines + (uint64_t)neighbour * 8u;
        cp_async16(&near_ring[warp][0][lane], p);
        cp_async16(&near_ring[warp][1][lane], p + 4);
      }
      cp_async_commit();
      look_ahead();
      cp_async_wait_but_last();  // the neighbour sectors (the last group is the k-mers of t3)
      if (second) {
        U32x8 sec;
        const uint4 lo4 = near_ring[warp][0][lane], hi4 = near_ring[warp][1][lane];
        sec.v[0] = lo4.x; sec.v[1] = lo4.y; sec.v[2] = lo4.z; sec.v[3] = lo4.w;
        sec.v[4] = hi4.x; sec.v[5] = hi4.y; sec.v[6] = hi4.z; sec.v[7] = hi4.w;
        uint32_t pos[4];
        const Sector s1 = classify_loaded<kTies>(ix, key, neighbour, sec, pol, pos);
        st = two_sector_second(ix, s0, s1, &b);
      }
    } else if (active && !done && st == 1) {  // round 2 in place
      uint32_t pos[4];
      const Sector s1 = classify_sector<kTies>(ix, key, neighbour, pol, pos);
      st = two_sector_second(ix, s0, s1, &b);
    }
    // phase 2 (kmer.cuh replay_plquery, in its pieces): the one loop in it runs with a warp-uniform trip count, so all
    // lanes are together again when rev[rank] is read
    const bool fin = active && !done && st == 0;
    ReplayState rs;
    rs.rank = kNoRank;
    rs.lo = 1u;  // lo > hi: not searching
    rs.hi = 0u;
    if (fin) {
      replay_windows(ix, pred, b, &rs);
      replay_search_jump(b, &rs);
    }
    while (__any_sync(0xffffffffu, rs.searching())) replay_search_step(b, &rs);  // :245; a no-op for the lanes that are done
    if (fin) {
      done = true;
      r = -1;                                                               // :246
      if (rs.rank != kNoRank) r = (long long)rev_at(ix, rs.rank, pol.sa);  // :247
    }
    // (stored a tile later instead, so that the read of rev[rank] has a tile to arrive: 9.6 against 5.5 ms, gpurun s17)
    if (active && done) store(i, x0, r);
    push(active && !done, x0, pred, i);
    while (stacked >= 32u) drain(32u);
    x0 = x1;
    x1 = x2;
    m1 = m2;
    pred0 = pred1;
    sec0 = sec1;
    t0 = t1;
    t1 = t2;
    t2 = t3;
    t3 = claim();
  }
  cp_async_wait_all();
  while (stacked) drain(stacked < 32u ? stacked : 32u);
}

// plQuery(s, kmer, length) for strings of any length (sapling_api.h:159): the literal replay with the gallop loops.
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
// :284 is on raw bytes against a pure-ACGT genome), so it is answered "no hit" without a query.
// Output per seed: ref_pos (0xFFFFFFFF = no verified hit), sa_pos, and the two hit counts as bytes (maxHits <= 255).
template <bool kTies>
__global__ void __launch_bounds__(kQueryThreads)
seed_kernel(const IndexView ix, const uint32_t* __restrict__ isa, const uint8_t* __restrict__ kflag,
            const char* __restrict__ reads, const uint64_t* __restrict__ off, size_t n_reads, uint32_t num_seeds,
            uint32_t maxHits, uint32_t* __restrict__ ref_pos, uint32_t* __restrict__ sa_pos,
            uint8_t* __restrict__ left, uint8_t* __restrict__ right) {
  const size_t total = n_reads * 2 * (size_t)num_seeds;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint32_t k = (uint32_t)ix.k;
  const L2Policies pol = make_policies(ix.hints);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const size_t r = t / (2 * (size_t)num_seeds);
    const uint32_t rem = (uint32_t)(t - r * 2 * (size_t)num_seeds);
    const uint32_t strand = rem / num_seeds, i = rem - strand * num_seeds;
    const uint64_t o = off[r], len = off[r + 1] - o;
    uint32_t hit = 0xFFFFFFFFu;
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
        const uint32_t pred = (uint32_t)clamp_prediction(ix, predict_rank(ix, x, pol.model));
        const long long a = answer_kmer_fast<kTies>(ix, x, pred, pol);
        if (a >= 0 && (uint64_t)a + k <= ix.n &&
            (load_bases_upto(ix.genome, (uint64_t)a, k) >> (64u - 2u * k)) == x) {
          hit = (uint32_t)a;
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
    left[t] = (uint8_t)lf;
    right[t] = (uint8_t)rt;
  }
}

__global__ void predict_kernel(const IndexView ix, const uint64_t* __restrict__ kmers, size_t nq,
                               uint64_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride)
    out[i] = predict_rank(ix, kmers[i], make_policies(0).model);
}

// k-mers that arrive as a dense little-endian bit stream, kmer_bits (< 64) bits each (the narrow upload formats of the
// host path: whole bytes, or exactly 2k bits) -> one 64-bit word each.  Two aligned 8-byte loads and a funnel shift per
// k-mer; the buffer is readable 8 bytes past its end.
__global__ void unpack_kmers_kernel(const uint64_t* __restrict__ raw, int kmer_bits, size_t nq, uint64_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint64_t mask = (1ull << kmer_bits) - 1ull;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
    const size_t bit = i * (size_t)kmer_bits;
    const unsigned sh = (unsigned)(bit & 63u);
    const uint64_t lo = __ldg(raw + (bit >> 6));
    uint64_t v = lo >> sh;
    if (sh + (unsigned)kmer_bits > 64u) v |= __ldg(raw + (bit >> 6) + 1) << (64u - sh);
    out[i] = v & mask;
  }
}

// rev[first .. first+count) out of the rank lines (Sapling::rev for callers that want it as an array)
__global__ void rev_extract_kernel(const IndexView ix, uint64_t first, uint64_t count, uint32_t* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const uint64_t r = first + i;
    out[i] = __ldg(ix.lines + (r >> 2) * 8u + 4u + (r & 3u));
  }
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
// query, i.e. the literal replay without any shortcut; the final unverified rev[lo + 1] (:136,:247) is a suffix-array
// read, not a getLcp call.  bench.py uses the mean as `probes_per_query` on every rank.
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
    Replay<false, KmerQuery, false> rp;
    rp.begin(pred);
    long long result;
    for (;;) {
      const bool final_read = rp.state == ST_FINAL;
      const bool done = rp.step(ix, q, pol, &result);
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
               unsigned long long* __restrict__ sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nthreads_work; i += stride) {
    uint64_t h = splitmix64(salt + i);
    for (int c = 0; c < chain; c++) {
      const uint64_t u = __umul64hi(h, nunits);
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

// Whether an entry of the rank lines can tie with a k-mer (the k-mer is longer than the entry's prefix).
static inline bool has_ties(const IndexView& ix) { return ix.k > ix.line_bases; }

const char* kmer_query_kernel_name(bool ordered) { return ordered ? "kmer_query_ordered_kernel" : "kmer_query_kernel"; }

// Resident blocks per SM the kernels are compiled for (register cap = 65536 / (256 * blocks)).  Measured defaults;
// `occupancy` (Tuning, capi.cu) overrides for A/B runs.
int kmer_query_blocks_per_sm(bool ordered, int occupancy) {
  if (ordered) return kOrderedBlocks;
  if (occupancy >= 3 && occupancy <= 6) return occupancy;
  return 4;
}

int launch_kmer_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                      int occupancy, cudaStream_t st) {
  if (nq == 0) return 0;
  const int bps = kmer_query_blocks_per_sm(false, occupancy);
  const int grid = query_grid(nq, bps * 2);
  const bool ties = has_ties(ix);
#define SB_LAUNCH(B)                                                                                                 \
  do {                                                                                                               \
    if (d_out32) {                                                                                                   \
      if (ties) kmer_query_kernel<B, true, uint32_t><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out32);      \
      else kmer_query_kernel<B, false, uint32_t><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out32);          \
    } else {                                                                                                         \
      if (ties) kmer_query_kernel<B, true, long long><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out);       \
      else kmer_query_kernel<B, false, long long><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out);           \
    }                                                                                                                \
  } while (0)
  switch (bps) {
    case 3: SB_LAUNCH(3); break;
    case 5: SB_LAUNCH(5); break;
    case 6: SB_LAUNCH(6); break;
    default: SB_LAUNCH(4); break;
  }
#undef SB_LAUNCH
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// The in-order kernel.  d_tiles is its zeroed tile counter.  Partitioned batch (partition.cu): d_slot is the slot array or
// slot_in_kmer_tag(), the results are slot words in d_res.  Batch in the caller's order (d_slot == nullptr): answers go to
// d_out (long long) or d_out32 (uint32_t), whichever is given.
template <int kOut>
static void launch_ordered(const IndexView& ix, const uint64_t* d_kmers, size_t nq, void* d_out, const uint16_t* d_slot,
                           unsigned long long* d_tiles, cudaStream_t st) {
  const bool ties = has_ties(ix), narrow = ix.narrow != nullptr;
  const int grid = query_grid(nq, kOrderedBlocks);
  if (ties && narrow) kmer_query_ordered_kernel<true, true, kOut><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);
  else if (ties) kmer_query_ordered_kernel<true, false, kOut><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);
  else if (narrow) kmer_query_ordered_kernel<false, true, kOut><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);
  else kmer_query_ordered_kernel<false, false, kOut><<<grid, kQueryThreads, 0, st>>>(ix, d_kmers, nq, d_out, d_slot, d_tiles);
}
int launch_kmer_query_ordered(const IndexView& ix, const uint64_t* d_part_kmers, size_t nq, long long* d_res,
                              const uint16_t* d_slot, unsigned long long* d_tiles, cudaStream_t st) {
  if (nq == 0) return 0;
  launch_ordered<kOutSlotWord>(ix, d_part_kmers, nq, d_res, d_slot, d_tiles, st);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_kmer_query_inorder(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                              unsigned long long* d_tiles, cudaStream_t st) {
  if (nq == 0) return 0;
  if (nq >= (1ull << 32)) { set_error("launch_kmer_query_inorder: nq=%zu out of range", nq); return -1; }
  SB_CUDA_CHECK(cudaMemsetAsync(d_tiles, 0, sizeof(unsigned long long), st));
  if (d_out32) launch_ordered<kOutU32>(ix, d_kmers, nq, d_out32, nullptr, d_tiles, st);
  else launch_ordered<kOutInt64>(ix, d_kmers, nq, d_out, nullptr, d_tiles, st);
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
                 const uint64_t* d_off, size_t n_reads, uint32_t num_seeds, uint32_t maxHits, uint32_t* d_ref_pos,
                 uint32_t* d_sa_pos, uint8_t* d_left, uint8_t* d_right, cudaStream_t st) {
  const size_t total = n_reads * 2 * (size_t)num_seeds;
  if (total == 0) return 0;
  if (has_ties(ix))
    seed_kernel<true><<<query_grid(total, 8), kQueryThreads, 0, st>>>(ix, d_isa, d_kflag, d_reads, d_off, n_reads, num_seeds,
                                                                      maxHits, d_ref_pos, d_sa_pos, d_left, d_right);
  else
    seed_kernel<false><<<query_grid(total, 8), kQueryThreads, 0, st>>>(ix, d_isa, d_kflag, d_reads, d_off, n_reads, num_seeds,
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

int launch_unpack_kmers(const void* d_packed, int kmer_bits, size_t nq, uint64_t* d_kmers, cudaStream_t st) {
  if (nq == 0) return 0;
  if (kmer_bits < 2 || kmer_bits > 63) { set_error("unpack_kmers: kmer_bits=%d out of range", kmer_bits); return -1; }
  unpack_kmers_kernel<<<query_grid(nq, 8), kQueryThreads, 0, st>>>(static_cast<const uint64_t*>(d_packed), kmer_bits, nq,
                                                                   d_kmers);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_rev_extract(const IndexView& ix, uint64_t first, uint64_t count, uint32_t* d_out, cudaStream_t st) {
  if (count == 0) return 0;
  rev_extract_kernel<<<query_grid(count, 8), kQueryThreads, 0, st>>>(ix, first, count, d_out);
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
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    const uint64_t salt = 0x9999ull + (uint64_t)r * work;
    cudaEventRecord(e0);
    if (gran == 16) gather2_kernel<16><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink);
    else if (gran == 32) gather2_kernel<32><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink);
    else if (gran == 64) gather2_kernel<64><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink);
    else gather2_kernel<128><<<grid, 256>>>(reinterpret_cast<const uint4*>(buf), nunits, work, chain, salt, sink);
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

// query.cuh -- device-side replay of Sapling::plQuery (reference sapling_api.h:159-248) with
// binarySearch (:133-153) and getLcp (:115-120) on the packed index.
//
// The reference's control flow is restated as a state machine whose every iteration performs
// exactly one probe (one SA load followed by one packed-genome compare).  All lanes of a warp
// therefore issue their dependent loads from the same instruction, whatever branch of plQuery
// they are in; only the (cheap, register-only) transition code diverges.
#pragma once

#include "common.cuh"

namespace sb {

enum ProbeState : int {
  ST_PRED = 0,  // probing rev[predicted]                      (:162-164)
  ST_R1,        // probing hi = min(n-1, predicted+mostOver)    (:171-174)
  ST_R2,        // probing hi = min(n-1, predicted+maxOver+1)   (:180-183)
  ST_RG,        // galloping right by maxOver (s.length() > k)  (:186-195)
  ST_L1,        // probing lo = max(0,(int)predicted-mostUnder) (:209-213)
  ST_L2,        // probing lo = max(0,(int)predicted-maxUnder-1)(:225-228)
  ST_LG,        // galloping left by maxUnder (s.length() > k)  (:231-240)
  ST_BS,        // binarySearch probing mid                     (:138-141)
  ST_FINAL,     // hi == lo+2: return rev[lo+1] unverified      (:136,:247)
  ST_SKIP       // verifying a run of skipped binarySearch steps (see pl_query_from)
};

struct ProbeResult {
  uint32_t lcp;   // getLcp result
  bool at_end;    // lcp + idx == n
  bool q_gt;      // s[lcp] > reference[idx+lcp]   (meaningful when !at_end)
  bool q_lt;      // s[lcp] < reference[idx+lcp]   (meaningful when !at_end)
};

// ---- query flavours ---------------------------------------------------------------------------

// A k-mer given as its kmerize() value: s = the k bases, s.length() == length == k <= 32.
struct KmerQuery {
  uint64_t q;  // bases left-aligned (base 0 in the top two bits)
  uint32_t k;
  __device__ __forceinline__ uint32_t slen() const { return k; }
  __device__ __forceinline__ uint32_t length() const { return k; }
  __device__ __forceinline__ uint64_t head() const { return q; }  // first 32 bases, left-aligned
  __device__ __forceinline__ ProbeResult probe(const IndexView& ix, uint64_t idx, uint32_t start,
                                               uint64_t pol) const {
    return compare(ix, idx, load_bases_upto_pol(ix.genome, idx, k, pol), start);
  }
  // g = the bases of the suffix at idx, left-aligned (at least k of them, or up to the end of the text)
  __device__ __forceinline__ ProbeResult compare(const IndexView& ix, uint64_t idx, uint64_t g, uint32_t start) const {
    uint64_t diff = q ^ g;
    if (start) diff &= (~0ull) >> (2u * start);  // getLcp trusts the first `start` characters
    const uint32_t m = diff ? ((uint32_t)__clzll((long long)diff) >> 1) : 32u;
    const uint64_t room = ix.n - idx;            // characters left in the text
    const uint32_t leff = room < (uint64_t)k ? (uint32_t)room : k;
    ProbeResult r;
    r.lcp = m < leff ? m : leff;
    r.at_end = (uint64_t)r.lcp == room;
    const uint32_t sh = 62u - 2u * (r.lcp & 31u);
    const uint32_t qb = (uint32_t)(q >> sh) & 3u, gb = (uint32_t)(g >> sh) & 3u;
    r.q_gt = qb > gb;
    r.q_lt = qb < gb;
    return r;
  }
};

// A string of slen bases (2-bit packed, 32 per word, left-aligned, zero padded) queried with the
// reference's separate `length` argument (plQuery(s, kmer, length), length <= slen).
struct StringQuery {
  const uint64_t* __restrict__ w;
  uint32_t slen_, length_;
  __device__ __forceinline__ ProbeResult compare(const IndexView&, uint64_t, uint64_t, uint32_t) const {
    return ProbeResult{};  // strings never use the inline-prefix path
  }
  __device__ __forceinline__ uint32_t slen() const { return slen_; }
  __device__ __forceinline__ uint32_t length() const { return length_; }
  __device__ __forceinline__ uint64_t head() const { return w[0]; }
  __device__ __forceinline__ ProbeResult probe(const IndexView& ix, uint64_t idx, uint32_t start,
                                               uint64_t pol) const {
    const uint64_t room = ix.n - idx;
    const uint32_t leff = room < (uint64_t)length_ ? (uint32_t)room : length_;
    uint32_t lcp = leff > start ? leff : start;
    for (uint32_t pos = start & ~31u; pos < leff; pos += 32u) {
      uint64_t diff = w[pos >> 5] ^ load_bases32_pol(ix.genome, idx + pos, pol);
      if (pos < start) diff &= (~0ull) >> (2u * (start - pos));
      if (diff) {
        const uint32_t m = pos + ((uint32_t)__clzll((long long)diff) >> 1);
        lcp = m < leff ? m : leff;
        break;
      }
    }
    ProbeResult r;
    r.lcp = lcp;
    r.at_end = (uint64_t)lcp == room;
    r.q_gt = r.q_lt = false;
    if (!r.at_end && lcp < slen_) {
      const uint32_t qb = (uint32_t)(w[lcp >> 5] >> (62u - 2u * (lcp & 31u))) & 3u;
      const uint64_t gi = idx + lcp;
      const uint32_t gb = (uint32_t)(ld_u64_pol(ix.genome + (gi >> 5), pol) >> (62u - 2u * (uint32_t)(gi & 31u))) & 3u;
      r.q_gt = qb > gb;
      r.q_lt = qb < gb;
    }
    return r;
  }
};

// ---- the replay --------------------------------------------------------------------------------

// Clamp of an out-of-range prediction.  The reference reads rev[] out of bounds there;
// defined here as: count it, use the last rank.
__device__ __forceinline__ uint64_t clamp_prediction(const IndexView& ix, uint64_t pred) {
  if (pred >= ix.n) {
    atomicAdd(ix.oob_counter, 1ull);
    pred = ix.n - 1;
  }
  return pred;
}

#ifdef SB_HOST_SIM
// tests/sim only: how often the long-window shortcut was tried / accepted
static unsigned long long g_sim_skip_tried = 0, g_sim_skip_ok = 0;
#define SB_SIM_COUNT(x) (++(x))
#else
#define SB_SIM_COUNT(x) ((void)0)
#endif

// ---- suffix-array readers -----------------------------------------------------------------------
// Every rev[r] read of the replay goes through one of these.

// plain gather: one 4-byte load per read
struct SaDirect {
  __device__ __forceinline__ uint32_t ld(const IndexView& ix, uint64_t r, uint64_t pol) const {
    return ld_u32_pol(ix.sa + r, pol);
  }
};

// Rank-line reader (layout: common.cuh IndexView).  The query's anchor line is the one that contains
// [predicted - min(mostUnder, 4), +16 ranks): with overlapping lines (packed_shift 3) every rank the typical replay
// touches is served by that ONE 128-byte DRAM line; ranks outside it are read from the line that starts at or
// below them.  The last sector read stays in registers (the replay often asks for neighbouring ranks in turn).
struct SaPacked {
  uint64_t abase;  // first rank of the anchor line
  uint64_t cur;    // first rank of the sector held in e (multiple of 4); ~0: none
  U32x8 e;
  __device__ __forceinline__ void anchor(const IndexView& ix, uint64_t pred) {
    const uint64_t back = (uint64_t)(ix.mostUnder < 4 ? ix.mostUnder : 4);
    const uint64_t lo = ix.packed_shift == 3 ? (pred > back ? pred - back : 0) : pred;
    abase = (lo >> ix.packed_shift) << ix.packed_shift;
    cur = ~0ull;
    // (Prefetching the other three sectors of the anchor line into L1 here was measured and rejected: prefetch.global.L1
    // is slower than the demand loads it saves, 2.6 -> 6.7 ms per 50 M queries at c2; gpurun r1z.)
  }
  // rev[r]; *g = the entry's leading bases left-aligned; *esc = compare against the packed genome instead
  __device__ __forceinline__ uint64_t get(const IndexView& ix, uint64_t r, uint64_t pol, uint64_t* g, bool* esc) {
    const uint64_t s4 = r & ~3ull;
    if (s4 != cur) {
      const uint64_t first = (r - abase < 16) ? abase : ((r >> ix.packed_shift) << ix.packed_shift);
      const uint64_t sector = (first >> ix.packed_shift) * 4 + ((r - first) >> 2);
      e = ld_u32x8_pol(ix.packed + sector * 8, pol);
      cur = s4;
    }
    const unsigned j = (unsigned)r & 3u;
    const uint64_t P0 = ((uint64_t)e.v[1] << 32) | e.v[0];
    const uint64_t D = ((uint64_t)e.v[3] << 32) | e.v[2];
    const uint64_t d = j ? ((D >> (kPackedDeltaBits * (j - 1))) & (uint64_t)kPackedEscape) : 0ull;
    *esc = j ? (d == (uint64_t)kPackedEscape) : ((D >> 63) != 0);
    *g = (P0 + d) << (64 - 2 * ix.packed_bases);
    const uint32_t a = (j & 1u) ? e.v[5] : e.v[4], b = (j & 1u) ? e.v[7] : e.v[6];
    return (uint64_t)((j & 2u) ? b : a);
  }
};

#ifndef SB_HOST_SIM
// Line-cached reader.  Almost every rank a query touches (predicted, predicted +- mostOver/mostUnder, the
// binary-search mids and the final lo+1) lies within a few entries of `predicted`, but the reads are
// separated by dependent genome probes (~1 us), by which time neither L1 nor L2 still holds the line
// (ncu, profiles/r1b: 3.3 SA sector fetches per query reach DRAM for 1.25 distinct lines).  So the thread
// fetches the aligned 64-byte line (16 ranks: one DRAM burst) around `predicted` ONCE, with cp.async
// (LDGSTS: global -> shared without staging registers, bypassing L1), and serves later reads from shared
// memory; ranks outside the line fall back to a global load.
// Shared-memory layout of one buffer: uint4 chunk c (0..3) of thread t at buf[c * kThreads + t].
template <int kThreads>
struct SaLine {
  uint4* buf;     // this thread's column: buf[c * kThreads], c = 0..3
  uint64_t base;  // first rank of the cached line (multiple of 16)
  __device__ __forceinline__ void issue(const IndexView& ix, uint64_t pred, uint64_t pol) {
    base = pred & ~15ull;
    const uint4* p = reinterpret_cast<const uint4*>(ix.sa + base);  // the SA allocation is padded to 16 entries
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf);
#pragma unroll
    for (int c = 0; c < 4; c++)
      asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + c * kThreads * 16),
                   "l"(p + c), "l"(pol)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // wait until at most kPending later groups are still in flight
  template <int kPending>
  __device__ __forceinline__ void wait() const {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
  }
  __device__ __forceinline__ uint32_t ld(const IndexView& ix, uint64_t r, uint64_t pol) const {
    if ((r & ~15ull) == base) {
      const unsigned j = (unsigned)(r & 15u);
      return reinterpret_cast<const uint32_t*>(buf + (j >> 2) * kThreads)[j & 3u];
    }
    return ld_u32_pol(ix.sa + r, pol);
  }
};
#endif

#ifndef SB_HOST_SIM
// Sector-cached reader: the aligned 32-byte sector (8 ranks) around `predicted` is fetched with ONE 256-bit load
// and kept in registers; the later reads of the replay (predicted +- mostOver/mostUnder, mids, final lo+1) that
// fall into it cost no memory request at all.  Measured motivation (profiles/r1_experiments.md): the kernel is
// bound by the number of L1-missing requests in flight per SM, and the plain reader spends 3.3 of its ~6.8
// requests per query re-reading this one sector.
struct SaSector {
  U32x8 e;
  uint64_t base;  // first rank of the cached sector (multiple of 8)
  __device__ __forceinline__ void fill(const IndexView& ix, uint64_t pred, uint64_t pol) {
    base = pred & ~7ull;
    e = ld_u32x8_pol(ix.sa + base, pol);  // the SA allocation is padded to whole lines
  }
  __device__ __forceinline__ uint32_t ld(const IndexView& ix, uint64_t r, uint64_t pol) const {
    if ((r & ~7ull) == base) {
      const unsigned j = (unsigned)r & 7u;
      const uint32_t a = (j & 1u) ? e.v[1] : e.v[0], b = (j & 1u) ? e.v[3] : e.v[2];
      const uint32_t c = (j & 1u) ? e.v[5] : e.v[4], d = (j & 1u) ? e.v[7] : e.v[6];
      const uint32_t ab = (j & 2u) ? b : a, cd = (j & 2u) ? d : c;
      return (j & 4u) ? cd : ab;
    }
    return ld_u32_pol(ix.sa + r, pol);
  }
};
#endif

// The replay from a known prediction.  kHaveFirst: idx0 = rev[pred] was already loaded by the caller
// (software-pipelined kernels issue that load one query ahead).
// kMode: 0 = {suffix array, packed genome}; 1 = inline-prefix entries (ExtEntry); 2 = rank lines (SaPacked).
//
// The replay of one query as an explicit state: begin() after the prediction, then step() once per probe until it
// returns true.  pl_query_from() below simply loops; the lane-refill kernel (query.cu) interleaves the steps of
// different queries on one lane so that a warp never waits for its slowest query.
template <bool kGallop, bool kHaveFirst, typename Query, typename Sa, bool kSkip = true, int kMode = 0>
struct Replay {
  uint64_t pred, lo, hi, r;
  uint32_t loLcp, hiLcp, lcp0, start;
  int state;

  __device__ __forceinline__ void begin(uint64_t predicted) {
    pred = predicted;
    lo = 0;
    hi = 0;
    r = predicted;
    loLcp = hiLcp = lcp0 = start = 0;
    state = ST_PRED;
  }

  // one probe; true when the query is answered (*result = the reference's return value)
  __device__ __forceinline__ bool step(const IndexView& ix, const Query& qy, const uint64_t idx0, const L2Policies& pol,
                                       Sa& sa, long long* result) {
    const uint64_t n = ix.n;
    const uint64_t nm1 = n - 1;
    const uint32_t slen = qy.slen(), length = qy.length();
    // (int)predicted of :209/:225 -- wraps negative for predicted >= 2^31
    const int32_t p32 = (int32_t)(uint32_t)pred;
    uint64_t idx, g = 0;
    bool use_genome = true;
    if constexpr (kMode == 1) {  // one 16-byte entry holds rev[r] and the bases to compare
      const uint4 e = ld_u32x4_pol(reinterpret_cast<const uint4*>(ix.ext + r), pol.sa);
      idx = e.x;
      g = ((uint64_t)e.w << 32) | e.z;
      use_genome = false;
    } else if constexpr (kMode == 2) {  // one 32-byte sector holds four such entries
      idx = sa.get(ix, r, pol.sa, &g, &use_genome);
      // an entry carries packed_bases bases: a longer query that agrees on all of them is decided by the genome
      if (!use_genome && (int)length > ix.packed_bases)
        use_genome = ((qy.head() ^ g) >> (64 - 2 * ix.packed_bases)) == 0;
    } else {
      idx = (kHaveFirst && state == ST_PRED) ? idx0 : (uint64_t)sa.ld(ix, r, pol.sa);
    }
    if (state == ST_FINAL) { *result = (long long)idx; return true; }
    const ProbeResult pr = use_genome ? qy.probe(ix, idx, start, pol.genome) : qy.compare(ix, idx, g, start);
    const bool small = pr.at_end || pr.q_gt;  // "suffix too small" test of :143,:167,:175,:214
    bool to_search = false;
    switch (state) {
      case ST_PRED:
        if (pr.lcp == length) { *result = (long long)idx; return true; }  // :164
        lcp0 = pr.lcp;
        if (small) {  // :167-172
          lo = pred;
          hi = pred + (uint64_t)(long long)ix.mostOver;
          if (hi > nm1) hi = nm1;
          r = hi;
          state = ST_R1;
        } else {  // :209-211
          if (ix.compat) {
            const int32_t v = (int32_t)((uint32_t)p32 - (uint32_t)ix.mostUnder);
            lo = (uint64_t)(v > 0 ? v : 0);
          } else {
            const uint64_t d = (uint64_t)(long long)ix.mostUnder;
            lo = pred > d ? pred - d : 0;
          }
          hi = pred;
          r = lo;
          state = ST_L1;
        }
        break;
      case ST_R1:
        if (pr.lcp == length) { *result = (long long)idx; return true; }  // :174
        if (small) {                                  // :175-181
          lo = hi;
          loLcp = pr.lcp;
          hi = pred + (uint64_t)(long long)ix.maxOver + 1;
          if (hi > nm1) hi = nm1;
          r = hi;
          state = ST_R2;
        } else {  // :199-204
          loLcp = lcp0;
          hiLcp = pr.lcp;
          to_search = true;
        }
        break;
      case ST_R2:
      case ST_RG:
        if (pr.lcp == (state == ST_R2 ? length : slen)) { *result = (long long)idx; return true; }  // :183 / :194
        // :184-196 (reference loops forever once hi is pinned at n-1; we stop there)
        if (kGallop && slen > (uint32_t)ix.k && !pr.at_end && pr.q_gt && hi != nm1) {
          lo = hi;
          loLcp = pr.lcp;
          hi += (uint64_t)(long long)ix.maxOver;
          if (hi > nm1) hi = nm1;
          r = hi;
          state = ST_RG;
        } else {
          hiLcp = pr.lcp;  // :197
          to_search = true;
        }
        break;
      case ST_L1:
        if (pr.lcp == slen) { *result = (long long)idx; return true; }  // :213
        if (small) {                                // :214-219
          hiLcp = lcp0;
          loLcp = pr.lcp;
          to_search = true;
          // Long-window shortcut.  For ranks >= 2^31 the reference's (int)predicted arithmetic collapses lo to 0
          //  and binarySearch(0, predicted) then walks ~31 mids m_1 = (lo+hi)>>1, m_{j+1} = (m_j+hi)>>1,
          // nearly all of which "go right" because the query sits just left of `predicted`.  Those mids are a pure
          // function of (lo, hi).  Take the last one still at least maxUnder+1 ranks left of hi and probe only it:
          // if that suffix is strictly smaller than the query and not a match then, the suffix array being sorted,
          // so was every earlier mid (none can have matched either), and because every carried LCP on this branch
          // is a true LCP the reference arrives at exactly lo = that mid, loLcp = this probe's LCP.  Otherwise
          // nothing is assumed and the literal chain is replayed from (lo, hi).
          if (kSkip) {
            const uint64_t guard = (uint64_t)(long long)ix.maxUnder + 1;
            if (hi - lo > 4 * guard + 64) {
              uint64_t cand = lo;
              for (;;) {
                const uint64_t m = (cand + hi) >> 1;
                if (m + guard > hi) break;
                cand = m;
              }
              if (cand != lo) {
                SB_SIM_COUNT(g_sim_skip_tried);
                r = cand;
                start = 0;
                state = ST_SKIP;
                to_search = false;
              }
            }
          }
        } else {  // :220-226
          hi = lo;
          hiLcp = pr.lcp;
          if (ix.compat) {
            const int32_t v = (int32_t)((uint32_t)p32 - (uint32_t)ix.maxUnder - 1u);
            lo = (uint64_t)(v > 0 ? v : 0);
          } else {
            const uint64_t d = (uint64_t)(long long)ix.maxUnder + 1;
            lo = pred > d ? pred - d : 0;
          }
          r = lo;
          state = ST_L2;
        }
        break;
      case ST_L2:
      case ST_LG:
        if (pr.lcp == slen) { *result = (long long)idx; return true; }  // :228 / :239
        // :229-241 (reference underflows lo below rank 0; we stop at 0)
        if (kGallop && slen > (uint32_t)ix.k && !pr.at_end && pr.q_lt && lo != 0) {
          hi = lo;
          hiLcp = pr.lcp;
          const uint64_t step = (uint64_t)(long long)ix.maxUnder;
          lo = lo > step ? lo - step : 0;
          r = lo;
          state = ST_LG;
        } else {
          loLcp = pr.lcp;  // :242
          to_search = true;
        }
        break;
      case ST_SKIP:
        if (pr.lcp != slen && small) {  // verified: every skipped step went right
          SB_SIM_COUNT(g_sim_skip_ok);
          lo = r;
          loLcp = pr.lcp;
        }
        to_search = true;
        break;
      default:  // ST_BS, probed mid == r  (:139-152)
        if (pr.lcp == slen) { *result = (long long)idx; return true; }  // :141 then :247 (rev[mid] == idx)
        if (lo + 1 >= hi) { *result = -1; return true; }                // :142
        if (small) {
          lo = r;
          loLcp = pr.lcp;
        } else {
          hi = r;
          hiLcp = pr.lcp;
        }
        to_search = true;
        break;
    }
    if (to_search) {  // top of binarySearch (:136-140)
      if (hi == lo + 2) {
        r = lo + 1;
        state = ST_FINAL;
      } else {
        r = (lo + hi) >> 1;
        start = loLcp < hiLcp ? loLcp : hiLcp;
        state = ST_BS;
      }
    }
    return false;
  }
};

template <bool kGallop, bool kHaveFirst, typename Query, typename Sa, bool kSkip = true, int kMode = 0>
__device__ __forceinline__ long long pl_query_from(const IndexView& ix, const Query& qy, const uint64_t pred,
                                                   const uint64_t idx0, const L2Policies& pol, Sa& sa) {
  Replay<kGallop, kHaveFirst, Query, Sa, kSkip, kMode> rp;
  rp.begin(pred);
  long long result;
  while (!rp.step(ix, qy, idx0, pol, sa, &result)) {
  }
  return result;
}

// ---- lean k-mer replay --------------------------------------------------------------------------------------------
// The same replay specialised for what the batch kernels actually answer: a k-mer (s.length() == length == k <= 32, so no
// gallop loops) against an index with n < 2^32 ranks.  Ranks, windows and LCPs are 32-bit, and the nine cases of
// Replay::step collapse into ONE straight-line update, because every probe after the first either moves the low bound
// (lo = r, loLcp = lcp) or the high bound (hi = r, hiLcp = lcp) of the pair binarySearch is finally called with:
//   R1 (r == hi)  small: lo = hi, loLcp = lcp, then widen hi (:175-181)      else: hiLcp = lcp, search   (:199-204; loLcp
//                                                                                   already holds the first probe's lcp)
//   L1 (r == lo)  small: loLcp = lcp, search (:214-219; hiLcp already holds it)  else: hi = lo, hiLcp = lcp, then widen lo
//   R2 (r == hi)  hiLcp = lcp, search (:197)         L2 (r == lo)  loLcp = lcp, search (:242)
//   BS            small: lo = mid, loLcp = lcp       else: hi = mid, hiLcp = lcp    (:143-152)
//   SKIP          verified: lo = cand, loLcp = lcp   else: nothing                  (long-window shortcut, see Replay)
// A warp whose lanes sit in different cases therefore executes one instruction stream instead of one per case
// (ncu, profiles/r1y: 15 of 32 lanes active per instruction, 43 warp instructions per query with Replay::step).
// "Suffix too small" (:143) needs no base extraction: with the first `start` bases masked out of both words, the first
// differing base decides the unsigned comparison of the words.
// Preconditions (checked by the launcher, which otherwise uses Replay): n <= 2^32 - 16, error bounds >= 0.
constexpr uint64_t kLeanMaxN = 0xFFFFFFF0ull;

struct SaPacked32 {
  uint32_t abase;  // first rank of the anchor line
  uint32_t cur;    // first rank of the sector held in e (a multiple of 4); 0xFFFFFFFF: none
  U32x8 e;
  __device__ __forceinline__ void anchor(const IndexView& ix, uint32_t pred) {
    const uint32_t back = (uint32_t)(ix.mostUnder < 4 ? ix.mostUnder : 4);
    const uint32_t lo = ix.packed_shift == 3 ? (pred > back ? pred - back : 0u) : pred;
    abase = (lo >> ix.packed_shift) << ix.packed_shift;
    cur = 0xFFFFFFFFu;
  }
  __device__ __forceinline__ uint32_t get(const IndexView& ix, uint32_t r, uint64_t pol, uint64_t* g, bool* esc) {
    const uint32_t s4 = r & ~3u;
    if (s4 != cur) {
      uint32_t sector;
      if (ix.packed_shift == 4) {  // tiling lines: sector s holds ranks 4s .. 4s+3 (uniform branch, hoisted out of the loop)
        sector = r >> 2;
      } else {
        const uint32_t first = (r - abase < 16u) ? abase : ((r >> ix.packed_shift) << ix.packed_shift);
        sector = ((first >> ix.packed_shift) << 2) + ((r - first) >> 2);
      }
      e = ld_u32x8_pol(ix.packed + (uint64_t)sector * 8u, pol);
      cur = s4;
    }
    const unsigned j = r & 3u;
    const uint64_t P0 = ((uint64_t)e.v[1] << 32) | e.v[0];
    const uint64_t D = ((uint64_t)e.v[3] << 32) | e.v[2];
    const uint64_t d = j ? ((D >> (kPackedDeltaBits * (j - 1))) & (uint64_t)kPackedEscape) : 0ull;
    *esc = j ? (d == (uint64_t)kPackedEscape) : ((D >> 63) != 0);
    *g = (P0 + d) << (64 - 2 * ix.packed_bases);
    return pos_of(j);
  }
  __device__ __forceinline__ uint32_t pos_of(unsigned j) const {
    const uint32_t a = (j & 1u) ? e.v[5] : e.v[4], b = (j & 1u) ? e.v[7] : e.v[6];
    return (j & 2u) ? b : a;
  }
  // rev[r] if the sector in registers holds it (binarySearch's unverified rev[lo + 1] usually falls into it)
  __device__ __forceinline__ bool cached_pos(uint32_t r, uint32_t* idx) const {
    if ((r & ~3u) != cur) return false;
    *idx = pos_of(r & 3u);
    return true;
  }
};

// Anchor line staged in shared memory -- MEASURED SLOWER, opt-in (SAPLING_B200_LINE_SMEM=1); kept because the result is
// instructive: four sector requests per query instead of ~2.5 cost more than the dependent round trips they remove
// (gpurun r2c: c2 2.5 -> 4.9 ms unpartitioned, 1.6 -> 2.1 ms partitioned; c3 partitioned 12.3 ms, stable, against
// 10.1-15.8 ms).  The idea was: SaPacked32 above fetches the sectors of the anchor line one probe at a time and
// counts on L2 to still hold the line for the later ones.  Once the batch is walked in order that stops being true in a
// bistable way (gpurun r2b, c3: the same kernel takes 10 or 15.6 ms per 250 M queries depending on how many warps are
// resident): with ~5 TB/s of line fills streaming through L2 a line survives about as long as one query lasts, and when
// it does not, every later probe of the query refetches 128 bytes from DRAM and the queries get slower still.  Here the
// thread requests all four sectors of its anchor line at once (four independent 256-bit loads: one DRAM line fill, one
// latency instead of up to four dependent ones), parks them in its own 144-byte shared-memory slot and answers every
// probe inside the line from there; only ranks outside the anchor line (the long left chains of SURVEY F5, windows that
// straddle two tiling lines) still go to global memory, through a one-sector register cache as before.
// Slot stride 144 = 128 + 16 bytes: the 16-byte accesses of 8 consecutive lanes then fall into 8 different bank groups.
constexpr int kLineSlotU4 = 9;  // uint4 per thread slot
struct SaLine32 {
  uint4* sm;       // this thread's slot
  uint32_t abase;  // first rank of the anchor line
  uint32_t cur;    // first rank of the sector held in e (outside the anchor line); 0xFFFFFFFF: none
  U32x8 e;
  __device__ __forceinline__ void anchor(const IndexView& ix, uint32_t pred, uint64_t pol) {
    const uint32_t back = (uint32_t)(ix.mostUnder < 4 ? ix.mostUnder : 4);
    const uint32_t lo = ix.packed_shift == 3 ? (pred > back ? pred - back : 0u) : pred;
    abase = (lo >> ix.packed_shift) << ix.packed_shift;
    cur = 0xFFFFFFFFu;
    const uint32_t* line = ix.packed + (uint64_t)(abase >> ix.packed_shift) * 32u;
    const U32x8 s0 = ld_u32x8_pol(line, pol), s1 = ld_u32x8_pol(line + 8, pol);
    const U32x8 s2 = ld_u32x8_pol(line + 16, pol), s3 = ld_u32x8_pol(line + 24, pol);
    sm[0] = make_uint4(s0.v[0], s0.v[1], s0.v[2], s0.v[3]);
    sm[1] = make_uint4(s0.v[4], s0.v[5], s0.v[6], s0.v[7]);
    sm[2] = make_uint4(s1.v[0], s1.v[1], s1.v[2], s1.v[3]);
    sm[3] = make_uint4(s1.v[4], s1.v[5], s1.v[6], s1.v[7]);
    sm[4] = make_uint4(s2.v[0], s2.v[1], s2.v[2], s2.v[3]);
    sm[5] = make_uint4(s2.v[4], s2.v[5], s2.v[6], s2.v[7]);
    sm[6] = make_uint4(s3.v[0], s3.v[1], s3.v[2], s3.v[3]);
    sm[7] = make_uint4(s3.v[4], s3.v[5], s3.v[6], s3.v[7]);
  }
  __device__ __forceinline__ uint32_t get(const IndexView& ix, uint32_t r, uint64_t pol, uint64_t* g, bool* esc) {
    const unsigned j = r & 3u;
    uint64_t P0, D;
    uint32_t pos;
    const uint32_t off = r - abase;
    if (off < 16u) {
      const uint4 h = sm[2u * (off >> 2)];
      P0 = ((uint64_t)h.y << 32) | h.x;
      D = ((uint64_t)h.w << 32) | h.z;
      pos = reinterpret_cast<const uint32_t*>(sm + 2u * (off >> 2) + 1u)[j];
    } else {
      const uint32_t s4 = r & ~3u;
      if (s4 != cur) {
        const uint32_t first = (r >> ix.packed_shift) << ix.packed_shift;
        const uint32_t sector = ((first >> ix.packed_shift) << 2) + ((r - first) >> 2);
        e = ld_u32x8_pol(ix.packed + (uint64_t)sector * 8u, pol);
        cur = s4;
      }
      P0 = ((uint64_t)e.v[1] << 32) | e.v[0];
      D = ((uint64_t)e.v[3] << 32) | e.v[2];
      const uint32_t a = (j & 1u) ? e.v[5] : e.v[4], b = (j & 1u) ? e.v[7] : e.v[6];
      pos = (j & 2u) ? b : a;
    }
    const uint64_t d = j ? ((D >> (kPackedDeltaBits * (j - 1))) & (uint64_t)kPackedEscape) : 0ull;
    *esc = j ? (d == (uint64_t)kPackedEscape) : ((D >> 63) != 0);
    *g = (P0 + d) << (64 - 2 * ix.packed_bases);
    return pos;
  }
  __device__ __forceinline__ bool cached_pos(uint32_t r, uint32_t* idx) const {
    const uint32_t off = r - abase;
    if (off >= 16u) return false;
    *idx = reinterpret_cast<const uint32_t*>(sm + 2u * (off >> 2) + 1u)[r & 3u];
    return true;
  }
};

struct SaSector32 {
  U32x8 e;
  uint32_t base;  // first rank of the cached sector (multiple of 8)
  __device__ __forceinline__ void fill(const IndexView& ix, uint32_t pred, uint64_t pol) {
    base = pred & ~7u;
    e = ld_u32x8_pol(ix.sa + base, pol);  // the SA allocation is padded to whole lines
  }
  __device__ __forceinline__ uint32_t ld(const IndexView& ix, uint32_t r, uint64_t pol) const {
    if ((r & ~7u) == base) {
      const unsigned j = r & 7u;
      const uint32_t a = (j & 1u) ? e.v[1] : e.v[0], b = (j & 1u) ? e.v[3] : e.v[2];
      const uint32_t c = (j & 1u) ? e.v[5] : e.v[4], d = (j & 1u) ? e.v[7] : e.v[6];
      const uint32_t ab = (j & 2u) ? b : a, cd = (j & 2u) ? d : c;
      return (j & 4u) ? cd : ab;
    }
    return ld_u32_pol(ix.sa + r, pol);
  }
  __device__ __forceinline__ bool cached_pos(uint32_t r, uint32_t* idx) const {
    if ((r & ~7u) != base) return false;
    const unsigned j = r & 7u;
    const uint32_t a = (j & 1u) ? e.v[1] : e.v[0], b = (j & 1u) ? e.v[3] : e.v[2];
    const uint32_t c = (j & 1u) ? e.v[5] : e.v[4], d = (j & 1u) ? e.v[7] : e.v[6];
    const uint32_t ab = (j & 2u) ? b : a, cd = (j & 2u) ? d : c;
    *idx = (j & 4u) ? cd : ab;
    return true;
  }
};

struct SaNone32 {  // kMode 1 reads ExtEntry directly
  __device__ __forceinline__ bool cached_pos(uint32_t, uint32_t*) const { return false; }
};

__device__ __forceinline__ uint32_t lean_uhadd(uint32_t a, uint32_t b) {  // floor((a + b) / 2) without overflow
#ifdef SB_HOST_SIM
  return (uint32_t)(((uint64_t)a + b) >> 1);
#else
  return __uhadd(a, b);
#endif
}
__device__ __forceinline__ uint32_t lean_add_clamped(uint32_t a, uint32_t b, uint32_t top) {  // min(a + b, top), a <= top
  const uint32_t t = a + b;
  return (t < a || t > top) ? top : t;
}

// q = the k bases left-aligned; pred < n.  kMode: 0 = {suffix-array sector, packed genome}, 1 = inline-prefix entries,
// 2 = rank lines.  Returns plQuery's answer (sapling_api.h:159-248).
// One probe of the lean replay and the transition it causes; true = answered (*result).  kKnown tells the compiler
// which states the call can be in, so that each call site carries only the transitions that can happen there -- the
// kernel is bound by instruction issue (profiles/r2f_*), not by memory:
//   S_PRED  the first probe, rev[predicted]
//   -2      R1 or L1: the second probe of every query (the mostOver / mostUnder bound)
//   -3      anything but PRED, R1, L1: the third probe (R2, L2, the long-window shortcut, or already binarySearch)
//   -4      BS or FINAL: every probe after the third (R2, L2 and SKIP all lead into binarySearch)
//   -1      anything (not used by kmer_replay32; kept for callers that resume a replay)
enum : int { S_PRED = 0, S_R1, S_R2, S_L1, S_L2, S_BS, S_FINAL, S_SKIP };
struct Lean32 {
  uint32_t lo, hi, r, loLcp, hiLcp, start;
  int state;
};
template <int kMode, bool kSkip, int kKnown, typename Sa>
__device__ __forceinline__ bool lean_step(const IndexView& ix, const uint64_t q, const uint32_t pred, const L2Policies& pol,
                                          Sa& sa, Lean32& s, long long* result) {
  constexpr bool may_pred = kKnown == -1 || kKnown == S_PRED;
  constexpr bool may_first = kKnown == -1 || kKnown == -2;                  // R1, L1
  constexpr bool may_second = kKnown == -1 || kKnown == -3;                 // R2, L2, SKIP
  constexpr bool may_search = kKnown == -1 || kKnown == -3 || kKnown == -4;  // BS, FINAL
  const uint32_t k = (uint32_t)ix.k;
  const uint32_t n32 = (uint32_t)ix.n, nm1 = n32 - 1u;
  const int state = kKnown >= 0 ? kKnown : s.state;
  uint32_t& lo = s.lo;
  uint32_t& hi = s.hi;
  uint32_t& r = s.r;
  uint32_t& loLcp = s.loLcp;
  uint32_t& hiLcp = s.hiLcp;
  uint32_t& start = s.start;
  // ---- one probe: rev[r] and the leading bases of that suffix ----------------------------------------------------
  uint32_t idx;
  uint64_t g;
  if constexpr (kMode == 1) {
    const uint4 e = ld_u32x4_pol(reinterpret_cast<const uint4*>(ix.ext + r), pol.sa);
    idx = e.x;
    g = ((uint64_t)e.w << 32) | e.z;
    if (may_search && state == S_FINAL) { *result = (long long)idx; return true; }
  } else if constexpr (kMode == 2) {
    bool esc;
    idx = sa.get(ix, r, pol.sa, &g, &esc);
    if (may_search && state == S_FINAL) { *result = (long long)idx; return true; }
    // an entry carries packed_bases bases: a longer k-mer that agrees on all of them is decided by the genome
    if (!esc && (int)k > ix.packed_bases) esc = ((q ^ g) >> (64 - 2 * ix.packed_bases)) == 0;
    if (esc) g = load_bases_upto_pol(ix.genome, (uint64_t)idx, k, pol.genome);
  } else {
    idx = sa.ld(ix, r, pol.sa);
    if (may_search && state == S_FINAL) { *result = (long long)idx; return true; }
    g = load_bases_upto_pol(ix.genome, (uint64_t)idx, k, pol.genome);
  }
  // ---- getLcp from `start` (:115-120) and the "suffix too small" test (:143) --------------------------------------
  const uint32_t st0 = (kKnown == S_PRED || kKnown == -2) ? 0u : start;  // the first two probes compare from base 0
  const uint64_t mask = (~0ull) >> (2u * st0);  // start < k <= 32
  const uint64_t qm = q & mask, gm = g & mask;
  const uint64_t diff = qm ^ gm;
  const uint32_t m = diff ? ((uint32_t)__clzll((long long)diff) >> 1) : 32u;
  const uint32_t room = n32 - idx;  // characters left in the text
  const uint32_t leff = room < k ? room : k;
  const uint32_t lcp = m < leff ? m : leff;
  const bool small = (lcp == room) || (qm > gm);
  const bool match = lcp == k;

  if (may_pred && state == S_PRED) {  // :162-172 / :209-211
    if (match) { *result = (long long)idx; return true; }
    if (small) {
      lo = pred;
      loLcp = lcp;
      hi = lean_add_clamped(pred, (uint32_t)ix.mostOver, nm1);
      r = hi;
      s.state = S_R1;
    } else {
      hi = pred;
      hiLcp = lcp;
      if (ix.compat) {  // (int)predicted - mostUnder (:209): wraps negative for predicted >= 2^31
        const int32_t v = (int32_t)(pred - (uint32_t)ix.mostUnder);
        lo = (uint32_t)(v > 0 ? v : 0);
      } else {
        const uint32_t d = (uint32_t)ix.mostUnder;
        lo = pred > d ? pred - d : 0u;
      }
      r = lo;
      s.state = S_L1;
    }
    return false;
  }
  const bool is_skip = may_second && state == S_SKIP;
  if (!is_skip && match) { *result = (long long)idx; return true; }                          // :174 :183 :213 :228 :141
  if (may_search && (kKnown == -4 || state == S_BS) && lo + 1u >= hi) { *result = -1; return true; }  // :142
  {
    const bool to_lo = (may_second && state == S_R2) ? false : ((may_second && state == S_L2) ? true : small);
    const bool upd = !(is_skip && (!small || match));  // unverified shortcut: assume nothing
    if (upd) {
      if (to_lo) {
        lo = r;
        loLcp = lcp;
      } else {
        hi = r;
        hiLcp = lcp;
      }
    }
  }
  if (may_first && state == S_R1 && small) {  // :180-181
    hi = lean_add_clamped(pred, (uint32_t)ix.maxOver + 1u, nm1);
    r = hi;
    s.state = S_R2;
    return false;
  }
  if (may_first && state == S_L1 && !small) {  // :225-226
    if (ix.compat) {
      const int32_t v = (int32_t)(pred - (uint32_t)ix.maxUnder - 1u);
      lo = (uint32_t)(v > 0 ? v : 0);
    } else {
      const uint32_t d = (uint32_t)ix.maxUnder + 1u;
      lo = pred > d ? pred - d : 0u;
    }
    r = lo;
    s.state = S_L2;
    return false;
  }
  if (kSkip && may_first && state == S_L1) {  // small: the long-window shortcut (see Replay::step)
    const uint32_t guard = (uint32_t)ix.maxUnder + 1u;
    if ((uint64_t)(hi - lo) > 4ull * guard + 64ull) {
      uint32_t cand = lo;
      for (;;) {
        const uint32_t mid = lean_uhadd(cand, hi);
        if (hi - mid < guard) break;
        cand = mid;
      }
      if (cand != lo) {
        SB_SIM_COUNT(g_sim_skip_tried);
        r = cand;
        start = 0;
        s.state = S_SKIP;
        return false;
      }
    }
  }
#ifdef SB_HOST_SIM
  if (is_skip && small && !match) SB_SIM_COUNT(g_sim_skip_ok);
#endif
  // top of binarySearch (:136-140)
  if (hi - lo == 2u) {
    r = lo + 1u;
    uint32_t pos;
    if (sa.cached_pos(r, &pos)) { *result = (long long)pos; return true; }  // rev[lo + 1] without another trip round the loop
    s.state = S_FINAL;
  } else {
    r = lean_uhadd(lo, hi);
    start = loLcp < hiLcp ? loLcp : hiLcp;
    s.state = S_BS;
  }
  return false;
}

template <int kMode, bool kSkip, typename Sa>
__device__ __forceinline__ long long kmer_replay32(const IndexView& ix, const uint64_t q, const uint32_t pred,
                                                   const L2Policies& pol, Sa& sa) {
  Lean32 s;
  s.lo = 0; s.hi = 0; s.r = pred; s.loLcp = 0; s.hiLcp = 0; s.start = 0; s.state = S_PRED;
  long long result = 0;
  if (lean_step<kMode, kSkip, S_PRED>(ix, q, pred, pol, sa, s, &result)) return result;  // probe 1: rev[predicted]
  if (lean_step<kMode, kSkip, -2>(ix, q, pred, pol, sa, s, &result)) return result;      // probe 2: the mostOver / mostUnder bound
  if (lean_step<kMode, kSkip, -3>(ix, q, pred, pol, sa, s, &result)) return result;      // probe 3: R2 / L2 / shortcut / search
  for (;;) {
    if (lean_step<kMode, kSkip, -4>(ix, q, pred, pol, sa, s, &result)) return result;    // binarySearch
  }
}

// The same replay cut in two for the kernel that parks unfinished queries (query.cu kmer_query_ordered_kernel, kMode 5):
// head = the first three probes, after which a query is answered or sits in binarySearch (state BS or FINAL);
// tail = the binarySearch loop, resumed from a Lean32 with nothing cached in registers.
template <int kMode, bool kSkip, typename Sa>
__device__ __forceinline__ bool kmer_replay32_head(const IndexView& ix, const uint64_t q, const uint32_t pred,
                                                   const L2Policies& pol, Sa& sa, Lean32& s, long long* result) {
  s.lo = 0; s.hi = 0; s.r = pred; s.loLcp = 0; s.hiLcp = 0; s.start = 0; s.state = S_PRED;
  if (lean_step<kMode, kSkip, S_PRED>(ix, q, pred, pol, sa, s, result)) return true;
  if (lean_step<kMode, kSkip, -2>(ix, q, pred, pol, sa, s, result)) return true;
  return lean_step<kMode, kSkip, -3>(ix, q, pred, pol, sa, s, result);
}
template <int kMode, bool kSkip, typename Sa>
__device__ __forceinline__ long long kmer_replay32_tail(const IndexView& ix, const uint64_t q, const L2Policies& pol, Sa& sa,
                                                        Lean32& s) {
  long long result = 0;
  for (;;) {
    if (lean_step<kMode, kSkip, -4>(ix, q, 0u, pol, sa, s, &result)) return result;  // predicted is not used in binarySearch
  }
}

// ---- flat k-mer replay on tiling rank lines --------------------------------------------------------------------------
// MEASURED SLOWER than kmer_replay32 (gpurun r2g, see launch_kmer_query) and therefore opt-in (SAPLING_B200_FLAT=1); kept
// because the reason is instructive and the variant is covered by the parity tests.
// kmer_replay32 once more, for the layout and the kernel the partitioned batch path runs (rank lines with
// packed_shift == 4, so sector = rank >> 2 and no anchor line), written so that the compiler has nothing to branch on.
// ncu of the in-order kernel (profiles/r2f_*): 77-79 % of the issue slots busy, 16-18 of 32 lanes active per
// instruction, a ~215-instruction loop body full of per-state branches -- the kernel is bound by instruction issue, and
// lanes in different states of the replay serialise.  Here the state is ONE-HOT, every transition is a select on
// predicates all lanes compute, the three ways out of the loop are one test, and only the two rare paths (an escaped
// entry that needs the genome, the long-window shortcut) remain real branches.  Same probes in the same order, so the
// same answers (tests: host simulation against the oracle, GPU variants against each other).
template <bool kSkip>
__device__ __forceinline__ long long kmer_replay_flat(const IndexView& ix, const uint64_t q, const uint32_t pred,
                                                      const L2Policies& pol) {
  enum : uint32_t { F_PRED = 1u, F_R1 = 2u, F_R2 = 4u, F_L1 = 8u, F_L2 = 16u, F_BS = 32u, F_FINAL = 64u, F_SKIP = 128u };
  const uint32_t k = (uint32_t)ix.k;
  const uint32_t n32 = (uint32_t)ix.n, nm1 = n32 - 1u;
  const unsigned gsh = 64u - 2u * (unsigned)ix.packed_bases;
  const bool tie_possible = (int)k > ix.packed_bases;  // an entry holds fewer bases than the k-mer: ties go to the genome
  uint32_t lo = 0, hi = 0, r = pred, loLcp = 0, hiLcp = 0, start = 0, st = F_PRED;
  uint32_t cur = 0xFFFFFFFFu;
  U32x8 e;
#pragma unroll
  for (int i = 0; i < 8; i++) e.v[i] = 0;
  for (;;) {
    // ---- one probe: rev[r] and the leading bases of that suffix, from the sector r >> 2 ---------------------------
    const uint32_t sec = r >> 2;
    if (sec != cur) {
      e = ld_u32x8_pol(ix.packed + (uint64_t)sec * 8u, pol.sa);
      cur = sec;
    }
    const unsigned j = r & 3u;
    const uint64_t P0 = ((uint64_t)e.v[1] << 32) | e.v[0];
    const uint64_t D = ((uint64_t)e.v[3] << 32) | e.v[2];
    // entry 0 reads as delta 0 with bit 63 as its escape flag; entries 1-3 are 21-bit fields, all ones = escape
    const uint64_t dj = (D >> ((kPackedDeltaBits * j - kPackedDeltaBits) & 63u)) & (uint64_t)kPackedEscape;
    const uint64_t d = j ? dj : 0ull;
    bool esc = j ? (dj == (uint64_t)kPackedEscape) : ((D >> 63) != 0);
    uint64_t g = (P0 + d) << gsh;
    const uint32_t ia = (j & 1u) ? e.v[5] : e.v[4], ib = (j & 1u) ? e.v[7] : e.v[6];
    const uint32_t idx = (j & 2u) ? ib : ia;
    const bool fin = (st & F_FINAL) != 0;
    if (tie_possible) esc = esc || (((q ^ g) >> gsh) == 0);
    if (esc && !fin) g = load_bases_upto_pol(ix.genome, (uint64_t)idx, k, pol.genome);  // rare
    // ---- getLcp from `start` (:115-120) and the "suffix too small" test (:143) ------------------------------------
    const uint64_t mask = (~0ull) >> (2u * start);  // start < k <= 32
    const uint64_t qm = q & mask, gm = g & mask;
    const uint64_t diff = qm ^ gm;
    const uint32_t m = diff ? ((uint32_t)__clzll((long long)diff) >> 1) : 32u;
    const uint32_t room = n32 - idx;  // characters left in the text
    const uint32_t leff = room < k ? room : k;
    const uint32_t lcp = m < leff ? m : leff;
    const bool small = (lcp == room) || (qm > gm);
    const bool match = lcp == k;
    // ---- the three ways out: rev[lo + 1] unverified (:137-139), a verified match (:141 :174 :183 :213 :228), and
    //      an empty search interval (:142) -- in this order of precedence
    const bool hit = fin || (match && !(st & F_SKIP));
    const bool miss = (st & F_BS) && (lo + 1u >= hi);
    if (hit || miss) return hit ? (long long)idx : -1ll;
    // ---- every probe moves one bound of the pair binarySearch is finally called with (see kmer_replay32) ---------
    const bool to_lo = (st & F_R2) ? false : ((st & F_L2) ? true : small);
    const bool upd = !((st & F_SKIP) && (!small || match));  // unverified shortcut: assume nothing
    const bool set_lo = upd && to_lo, set_hi = upd && !to_lo;
    lo = set_lo ? r : lo;
    loLcp = set_lo ? lcp : loLcp;
    hi = set_hi ? r : hi;
    hiLcp = set_hi ? lcp : hiLcp;
    // ---- what to probe next ----------------------------------------------------------------------------------------
    const bool first = (st & F_PRED) != 0;
    const bool widen_hi = small && (st & (F_PRED | F_R1));    // :165-166 (mostOver), :180-181 (maxOver + 1)
    const bool widen_lo = !small && (st & (F_PRED | F_L1));   // :209-210 (mostUnder), :225-226 (maxUnder + 1)
    const uint32_t d_over = first ? (uint32_t)ix.mostOver : (uint32_t)ix.maxOver + 1u;
    const uint32_t d_under = first ? (uint32_t)ix.mostUnder : (uint32_t)ix.maxUnder + 1u;
    const uint32_t up = lean_add_clamped(pred, d_over, nm1);
    // (int)predicted - d (:209 :225) wraps negative for predicted >= 2^31 in the reference: compat keeps that
    const int32_t dv = (int32_t)(pred - d_under);
    const uint32_t down = ix.compat ? (uint32_t)(dv > 0 ? dv : 0) : (pred > d_under ? pred - d_under : 0u);
    if (kSkip && (st & F_L1) && small) {  // the long-window shortcut (see Replay::step); rare unless ranks >= 2^31
      const uint32_t guard = (uint32_t)ix.maxUnder + 1u;
      if ((uint64_t)(hi - lo) > 4ull * guard + 64ull) {
        uint32_t cand = lo;
        for (;;) {
          const uint32_t mid = lean_uhadd(cand, hi);
          if (hi - mid < guard) break;
          cand = mid;
        }
        if (cand != lo) {
          SB_SIM_COUNT(g_sim_skip_tried);
          r = cand;
          start = 0;
          st = F_SKIP;
          continue;
        }
      }
    }
#ifdef SB_HOST_SIM
    if ((st & F_SKIP) && small && !match) SB_SIM_COUNT(g_sim_skip_ok);
#endif
    hi = widen_hi ? up : hi;
    lo = widen_lo ? down : lo;
    const bool two = (hi - lo) == 2u;  // top of binarySearch (:136-140)
    const uint32_t r_bs = two ? lo + 1u : lean_uhadd(lo, hi);
    const uint32_t st_bs = two ? (uint32_t)F_FINAL : (uint32_t)F_BS;
    const uint32_t lcp_min = loLcp < hiLcp ? loLcp : hiLcp;
    r = widen_hi ? hi : (widen_lo ? lo : r_bs);
    // PRED -> R1 -> R2 is a shift; PRED -> L1 -> L2 is 8 then a shift
    st = widen_hi ? (st << 1) : (widen_lo ? (first ? (uint32_t)F_L1 : (uint32_t)F_L2) : st_bs);
    start = (widen_hi || widen_lo || two) ? start : lcp_min;
  }
}

// whether kmer_replay32 may answer queries on this index
__host__ __device__ inline bool lean_eligible(const IndexView& ix) {
  return ix.n <= kLeanMaxN && ix.k >= 1 && ix.k <= 32 && ix.maxOver >= 0 && ix.maxUnder >= 0 && ix.mostOver >= 0 &&
         ix.mostUnder >= 0;
}

template <bool kGallop, typename Query>
__device__ __forceinline__ long long pl_query(const IndexView& ix, const Query& qy, uint64_t kmer) {
  const L2Policies pol = make_policies(ix.hints);
  const uint64_t pred = clamp_prediction(ix, predict_rank(ix, kmer, pol.model));  // :161
  SaDirect sa;
  return pl_query_from<kGallop, false>(ix, qy, pred, 0, pol, sa);
}

}  // namespace sb

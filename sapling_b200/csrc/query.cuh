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

// rev[r] from the rank lines (sector r >> 2, position slot r & 3): one 4-byte load per read
struct SaDirect {
  __device__ __forceinline__ uint32_t ld(const IndexView& ix, uint64_t r, uint64_t pol) const {
    return ld_u32_pol(ix.lines + (r >> 2) * 8u + 4u + (r & 3u), pol);
  }
};

// The literal replay from a known prediction, for queries of any length (plQuery(s, kmer, length) with s.length() != k
// takes the gallop loops of :184-196,:229-241 and carries LCPs that matter): begin() after the prediction, then step()
// once per probe until it returns true.  The batch k-mer path does not use it (kmer.cuh); the string-query kernel and the
// probe counter of bench.py do.  rev[] is read from the rank lines, the compare is against the packed genome.
// kSkip: allow the verified long-window shortcut (below); false = the reference's probe sequence exactly.
template <bool kGallop, typename Query, bool kSkip = true>
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
  __device__ __forceinline__ bool step(const IndexView& ix, const Query& qy, const L2Policies& pol, long long* result) {
    const uint64_t n = ix.n;
    const uint64_t nm1 = n - 1;
    const uint32_t slen = qy.slen(), length = qy.length();
    // (int)predicted of :209/:225 -- wraps negative for predicted >= 2^31
    const int32_t p32 = (int32_t)(uint32_t)pred;
    const uint64_t idx = (uint64_t)SaDirect().ld(ix, r, pol.sa);
    if (state == ST_FINAL) { *result = (long long)idx; return true; }
    const ProbeResult pr = qy.probe(ix, idx, start, pol.genome);
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

template <bool kGallop, typename Query, bool kSkip = true>
__device__ __forceinline__ long long pl_query_from(const IndexView& ix, const Query& qy, const uint64_t pred,
                                                   const L2Policies& pol) {
  Replay<kGallop, Query, kSkip> rp;
  rp.begin(pred);
  long long result;
  while (!rp.step(ix, qy, pol, &result)) {
  }
  return result;
}

template <bool kGallop, typename Query>
__device__ __forceinline__ long long pl_query(const IndexView& ix, const Query& qy, uint64_t kmer) {
  const L2Policies pol = make_policies(ix.hints);
  const uint64_t pred = clamp_prediction(ix, predict_rank(ix, kmer, pol.model));  // :161
  return pl_query_from<kGallop>(ix, qy, pred, pol);
}

}  // namespace sb

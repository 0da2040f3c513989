// kmer.cuh -- the k-mer query path: plQuery(unpack(kmer), kmer, k) (reference sapling_api.h:159-248) answered from the
// rank lines in two separate phases, memory first, control flow second.
//
// Why this is the reference's answer.  For a k-mer (s.length() == length == k) every decision plQuery and binarySearch
// (:133-153) take is a function of how the probed suffix compares with the query -- smaller ("suffix too small", :143,
// :167: the first differing base, or the text ending first), equal on all k bases (:141, :164), or larger -- and the
// suffix array is sorted.  So there are two ranks lb <= ub with
//       rank r < lb  : smaller          lb <= r < ub : match          r >= ub : larger
// and the whole probe sequence of the reference, including the ranks >= 2^31 arithmetic of :209,:225 and the
// unverified rev[lo + 1] of :136, can be replayed in registers once lb and ub are known.  The carried LCPs (:140: getLcp
// starts at min(loLcp, hiLcp) without re-checking earlier bases) never change an outcome for a k-mer: every LCP the
// reference carries was computed from base 0 or from such a start, the suffixes at lo < mid < hi are sorted, hence
// LCP(query, suffix[mid]) >= min(LCP(query, suffix[lo]), LCP(query, suffix[hi])) whichever side of them the query lies
// on, so getLcp returns the true LCP and the comparison is the true one.  (Proved by the differential tests: the host
// build of this file against the oracle on every fixture genome, and the GPU against the unmodified reference at 3.1 Gbp.)
//
// Phase 1 (memory): find lb and ub by SECTOR -- one 32-byte sector of a rank line classifies four ranks at once with
// integer compares in delta space; the sector of the predicted rank settles most queries, its neighbour nearly all the
// rest, a gallop + bisection over sectors the tail.  The reference's probe chain (2.3 dependent probes per query at 100 Mbp,
// 4.3 at 3.1 Gbp where a third of the left-going queries walk ~31 binary-search steps) is not followed in memory at all.
// Phase 2 (registers): the literal replay of :159-248 over {lb, ub}; long runs of binary-search steps that all go right are
// jumped in closed form.  Phase 3: rev[answer rank] -- usually in a sector still in registers.
#pragma once

#include "common.cuh"

namespace sb {

#ifdef SB_HOST_SIM
static unsigned long long g_sim_sector_loads = 0, g_sim_slow_entries = 0, g_sim_final_loads = 0;
#define SB_SIM_ADD(x, v) ((x) += (v))
#else
#define SB_SIM_ADD(x, v) ((void)0)
#endif

// A meeting point for the lanes of a warp.  The value passes through an empty asm statement, so the compiler cannot tell
// which way a lane came from and cannot thread the branches in front of it past it: the paths that lead here join here
// (BSYNC) and the code behind it runs once for all of them, not once per path.
#define SB_MEET(v32) asm volatile("" : "+r"(v32))

// What a lane needs to classify sectors for one query.
struct KmerKey {
  uint64_t q;        // the k bases left-aligned (base 0 in the top two bits)
  uint64_t qlo;      // smallest / largest line_bases-base integer whose first min(k, line_bases) bases are the query's
  uint64_t qhi;
};
template <bool kTies>
__device__ __forceinline__ KmerKey make_key(const IndexView& ix, uint64_t x) {
  KmerKey key;
  const int k = ix.k, b = ix.line_bases;
  key.q = x << (64 - 2 * k);
  if (!kTies) {  // k <= line_bases: pad the k-mer with the smallest / largest bases
    key.qlo = x << (2 * (b - k));
    key.qhi = key.qlo | ((1ull << (2 * (b - k))) - 1ull);
  } else {  // an entry holds fewer bases than the k-mer: equality on them is a tie the genome decides (see classify)
    key.qlo = key.qhi = x >> (2 * (k - b));
  }
  return key;
}

// One classified sector: ranks 4s .. 4s+3.  c entries are smaller than the query, the next m match, the rest are larger
// (ranks past the end of the suffix array count as larger).
struct Sector {
  uint32_t s;  // sector number
  uint32_t c, m;
};

// Full compare of the query with the suffix at text position pos (escaped entries, ties): the rules of getLcp (:115-120)
// and of the "suffix too small" test (:143,:167) -- the text ending inside the k-mer makes the suffix the smaller one.
__device__ __forceinline__ void compare_with_genome(const IndexView& ix, const KmerKey& key, uint32_t pos, uint64_t pol,
                                                    bool* small, bool* match) {
  const uint32_t k = (uint32_t)ix.k;
  const uint64_t g = load_bases_upto_pol(ix.genome, (uint64_t)pos, k, pol);
  const uint64_t diff = key.q ^ g;
  const uint32_t m = diff ? ((uint32_t)__clzll((long long)diff) >> 1) : 32u;
  const uint64_t room = ix.n - (uint64_t)pos;
  const uint32_t leff = room < (uint64_t)k ? (uint32_t)room : k;
  const uint32_t lcp = m < leff ? m : leff;
  *match = lcp == k;
  *small = !*match && ((uint64_t)lcp == room || key.q > g);
  SB_SIM_ADD(g_sim_slow_entries, 1);
}

// Classify sector s (its 32 bytes in e); pos[] receives its four text positions.  Fast path (no escaped entry, no tie): the
// entries are P0 + d_j with 21-bit deltas, so
//   smaller  <=>  d_j < qlo - P0         match  <=>  qlo - P0 <= d_j <= qhi - P0
// with both bounds clamped into the delta range: four 32-bit compares each.
template <bool kTies>
__device__ __forceinline__ Sector classify_loaded(const IndexView& ix, const KmerKey& key, uint32_t s, const U32x8& e,
                                                  const L2Policies& pol, uint32_t pos[4]) {
  const uint64_t P0 = ((uint64_t)e.v[1] << 32) | e.v[0];
  const uint32_t dlo = e.v[2], dhi = e.v[3];
  const uint32_t d1 = dlo & kPackedEscape;
  const uint32_t d2 = ((dlo >> 21) | (dhi << 11)) & kPackedEscape;
  const uint32_t d3 = (dhi >> 10) & kPackedEscape;
  const long long a = (long long)(key.qlo - P0), b = (long long)(key.qhi - P0);  // line_bases <= 31: no overflow
  const int32_t a32 = a < 0 ? 0 : (a > (long long)kPackedEscape ? (int32_t)kPackedEscape : (int32_t)a);
  const int32_t b32 = b < 0 ? -1 : (b > (long long)kPackedEscape ? (int32_t)kPackedEscape : (int32_t)b);
  bool sm0 = 0 < a32, sm1 = (int32_t)d1 < a32, sm2 = (int32_t)d2 < a32, sm3 = (int32_t)d3 < a32;
  bool ma0 = !sm0 && 0 <= b32, ma1 = !sm1 && (int32_t)d1 <= b32, ma2 = !sm2 && (int32_t)d2 <= b32,
       ma3 = !sm3 && (int32_t)d3 <= b32;
  // entries the deltas cannot decide: escapes (delta overflow, suffix within 32 bases of the end of the text, padding past
  // the last rank) and, when the k-mer is longer than an entry's prefix, entries equal to the query on that prefix
  const bool e0 = (dhi >> 31) != 0, e1 = d1 == kPackedEscape, e2 = d2 == kPackedEscape, e3 = d3 == kPackedEscape;
  bool n0 = e0, n1 = e1, n2 = e2, n3 = e3;
  if (kTies) { n0 |= ma0; n1 |= ma1; n2 |= ma2; n3 |= ma3; }
  if (n0 | n1 | n2 | n3) {  // rare: decided by the packed genome, entry by entry
    const uint64_t r0 = (uint64_t)s * 4u;
    if (n0) { if (r0 < ix.n) compare_with_genome(ix, key, e.v[4], pol.genome, &sm0, &ma0); else sm0 = ma0 = false; }
    if (n1) { if (r0 + 1 < ix.n) compare_with_genome(ix, key, e.v[5], pol.genome, &sm1, &ma1); else sm1 = ma1 = false; }
    if (n2) { if (r0 + 2 < ix.n) compare_with_genome(ix, key, e.v[6], pol.genome, &sm2, &ma2); else sm2 = ma2 = false; }
    if (n3) { if (r0 + 3 < ix.n) compare_with_genome(ix, key, e.v[7], pol.genome, &sm3, &ma3); else sm3 = ma3 = false; }
  }
  Sector out;
  out.s = s;
  out.c = (uint32_t)sm0 + (uint32_t)sm1 + (uint32_t)sm2 + (uint32_t)sm3;
  out.m = (uint32_t)ma0 + (uint32_t)ma1 + (uint32_t)ma2 + (uint32_t)ma3;
  pos[0] = e.v[4]; pos[1] = e.v[5]; pos[2] = e.v[6]; pos[3] = e.v[7];
  return out;
}
__device__ __forceinline__ U32x8 load_sector(const IndexView& ix, uint32_t s, const L2Policies& pol) {
  SB_SIM_ADD(g_sim_sector_loads, 1);
  return ld_u32x8_pol(ix.lines + (uint64_t)s * 8u, pol.sa);
}
template <bool kTies>
__device__ __forceinline__ Sector classify_sector(const IndexView& ix, const KmerKey& key, uint32_t s, const L2Policies& pol,
                                                  uint32_t pos[4]) {
  return classify_loaded<kTies>(ix, key, s, load_sector(ix, s, pol), pol, pos);
}

struct Bounds {
  uint32_t lb, ub;  // ranks [0, lb) smaller than the query, [lb, ub) match, [ub, n) larger
};

// Phase 1: lb and ub by sector.  Both are the same search over sectors for the boundary of a monotone property --
//   mode 0: "every entry of the sector is smaller than the query"   (true for sectors left of lb's, false from it on)
//   mode 1: "every valid entry of the sector matches"               (true from lb's sector to the one before ub's)
// -- written as a state that is FED one classified sector at a time and answers with the next sector it wants, so that
// there is one classification site whatever a lane is doing (looking left or right, galloping or bisecting, extending the
// match run), and so that the kernels can re-pack the unfinished queries of a block between two sector loads.
// State: `yes` = the last sector known to have the property (-1: the virtual sector before the array), `no` = the first
// sector known not to have it (sector last_s + 1 with nothing in it: the virtual sector after the array).  While one of the
// two is virtual the search gallops away from the other (1, 2, 4 ... sectors), then it bisects.  The sector of the
// predicted rank (`first`) is remembered: the run of matches often ends in it.
struct Search {
  int32_t yes;
  uint32_t no_s, no_c, no_m;
  uint32_t first_c, first_m;  // the sector pred >> 2 (fed first)
  uint32_t step_log2;
  uint32_t mode;
  uint32_t lb;
  uint32_t t;  // the sector to classify next

  __device__ __forceinline__ void begin(const IndexView& ix, uint32_t pred) {
    yes = -1;
    no_s = (((uint32_t)ix.n - 1u) >> 2) + 1u;
    no_c = no_m = first_c = first_m = 0;
    step_log2 = 0;
    mode = 0;
    lb = 0;
    t = pred >> 2;
  }
  // x = the classification of sector t.  true: *out is final.  false: classify sector t (updated) and feed again.
  // Single exit and select-style updates on purpose: the lanes of a warp are in different cases here (looking left, right,
  // bisecting, extending a run), and a branch per case would keep them apart -- through the NEXT round's classification
  // as well, because the branches only reconverge where the paths that are done rejoin (ncu s3: the second round ran twice
  // per tile with 7 lanes each).
  // kFresh: the state is the one begin() left (the first sector of a query) -- told to the compiler, which then folds
  // most of the update away.
  template <bool kFresh = false>
  __device__ __forceinline__ bool feed(const IndexView& ix, uint32_t pred, const Sector& x, bool is_first, Bounds* out) {
    const uint32_t n32 = (uint32_t)ix.n;
    const int32_t last_s = (int32_t)((n32 - 1u) >> 2);
    const uint32_t last_valid = n32 - 4u * (uint32_t)last_s;
    const uint32_t valid = (int32_t)x.s == last_s ? last_valid : 4u;
    if (kFresh) {
      yes = -1;
      no_s = (uint32_t)(last_s + 1);
      no_c = no_m = 0;
      step_log2 = 0;
      mode = 0;
      is_first = true;
    }
    first_c = is_first ? x.c : first_c;
    first_m = is_first ? x.m : first_m;
    const bool has = mode ? (x.m == valid) : (x.c == 4u);
    yes = has ? (int32_t)x.s : yes;
    no_s = has ? no_s : x.s;
    no_c = has ? no_c : x.c;
    no_m = has ? no_m : x.m;
    bool done = false;
    // mode 0 ends when the boundary sector is known: lb lies in sector no_s (or is n: nothing but smaller suffixes)
    const bool lb_known = mode == 0 && (yes + 1 == (int32_t)no_s || ((int32_t)no_s <= last_s && no_c > 0u));
    if (lb_known) {
      const uint32_t raw = 4u * no_s + no_c;
      lb = raw < n32 ? raw : n32;
      const uint32_t nv = (int32_t)no_s == last_s ? last_valid : 4u;
      const bool ends_inside = (int32_t)no_s >= last_s || no_m == 0u || no_c + no_m < nv;
      out->lb = lb;
      out->ub = lb + no_m;
      done = ends_inside;
      // otherwise the run of matches reaches the end of the sector: search on for the first sector that is not all
      // matches (mode 1); the first sector, if it lies to the right, is already classified
      const int32_t fs = (int32_t)(pred >> 2);
      const uint32_t fv = fs == last_s ? last_valid : 4u;
      const bool use_first = fs > (int32_t)no_s;
      const bool first_all = first_m == fv;
      mode = ends_inside ? 0u : 1u;
      yes = (use_first && first_all) ? fs : (int32_t)no_s;
      step_log2 = 0;
      const bool first_is_no = use_first && !first_all;
      no_s = first_is_no ? (uint32_t)fs : (uint32_t)(last_s + 1);
      no_c = first_is_no ? first_c : 0u;
      no_m = first_is_no ? first_m : 0u;
    }
    if (!done && mode == 1u && yes + 1 == (int32_t)no_s) {
      const uint32_t raw = 4u * no_s + no_m;
      out->lb = lb;
      out->ub = raw < n32 ? raw : n32;
      done = true;
    }
    // the next sector to classify (computed for every lane; meaningless once done)
    const bool go_right = (int32_t)no_s > last_s;  // nothing known to the right yet: gallop right
    const bool go_left = !go_right && yes < 0;     // nothing known to the left yet: gallop left
    const uint32_t step = 1u << step_log2;
    const uint32_t cand_r = (uint32_t)yes + step < (uint32_t)last_s ? (uint32_t)yes + step : (uint32_t)last_s;
    const uint32_t cand_l = no_s > step ? no_s - step : 0u;
    const uint32_t cand_b = (uint32_t)yes + ((no_s - (uint32_t)yes) >> 1);
    t = go_right ? cand_r : (go_left ? cand_l : cand_b);
    step_log2 += (go_right || go_left) ? 1u : 0u;
    return done;
  }
  // the part of the state that is not recomputable, in four words (for the kernels' shared-memory queues)
  __device__ __forceinline__ uint4 pack() const {
    const uint32_t bits = no_c | (no_m << 3) | (first_c << 6) | (first_m << 9) | (step_log2 << 12) | (mode << 18);
    return make_uint4((uint32_t)yes, no_s, lb, bits);
  }
  __device__ __forceinline__ void unpack(const uint4& w, uint32_t next_t) {
    yes = (int32_t)w.x; no_s = w.y; lb = w.z;
    no_c = w.w & 7u; no_m = (w.w >> 3) & 7u; first_c = (w.w >> 6) & 7u; first_m = (w.w >> 9) & 7u;
    step_log2 = (w.w >> 12) & 63u; mode = (w.w >> 18) & 1u;
    t = next_t;
  }
};

// The common case of phase 1 without the general search: the sector of the predicted rank, and at most ONE neighbour.
// A third of the queries match at the predicted rank; for most of the others the boundary lb lies in that sector or in
// the next one to the side the classification points to, and the run of matches ends there too.  Everything else (errors
// beyond a sector, long runs of equal k-mers) is left to Search.
//   two_sector_first : 0 = *b is final; 1 = classify sector *neighbour and call two_sector_second
//   two_sector_second: 0 = *b is final; 2 = not decided by these two sectors
__device__ __forceinline__ int two_sector_first(const IndexView& ix, const Sector& x0, Bounds* b, uint32_t* neighbour) {
  const uint32_t n32 = (uint32_t)ix.n;
  const uint32_t last_s = (n32 - 1u) >> 2;
  const uint32_t v0 = x0.s == last_s ? n32 - 4u * last_s : 4u;
  const bool all_small = x0.c == 4u;                     // lb lies further right
  const bool none_small = x0.c == 0u && x0.s != 0u;      // lb may lie further left
  const bool is_last = x0.s == last_s;
  const uint32_t lb = all_small ? n32 : 4u * x0.s + x0.c;  // (all small in the last sector: every rank is smaller)
  const bool run_ends = x0.c + x0.m < v0 || is_last;
  b->lb = lb;
  b->ub = all_small ? n32 : lb + x0.m;
  *neighbour = none_small ? x0.s - 1u : x0.s + 1u;
  const bool resolved = all_small ? is_last : (!none_small && run_ends);
  return resolved ? 0 : 1;
}
__device__ __forceinline__ int two_sector_second(const IndexView& ix, const Sector& x0, const Sector& x1, Bounds* b) {
  // Three cases, all computed and then selected (the lanes of a warp are spread over them):
  //   right neighbour, s0 all smaller: lb and the run of matches are looked for in s1
  //   right neighbour, lb = 4 s0 + c0 known and the matches ran to the end of s0: where do they stop?
  //   left neighbour (s0 has no smaller entry): lb is looked for in s1, the matches may go on into s0
  const uint32_t n32 = (uint32_t)ix.n;
  const uint32_t last_s = (n32 - 1u) >> 2;
  const uint32_t v0 = x0.s == last_s ? n32 - 4u * last_s : 4u;
  const uint32_t v1 = x1.s == last_s ? n32 - 4u * last_s : 4u;
  const bool right = x1.s > x0.s, want_lb = x0.c == 4u;
  const bool x1_last = x1.s == last_s;
  const uint32_t lb1 = 4u * x1.s + x1.c;
  const bool run_ends = x1.c + x1.m < v1 || x1_last;
  // right, looking for lb
  const bool all_small = x1.c == 4u;
  const uint32_t lb_r = all_small ? n32 : lb1;
  const uint32_t ub_r = all_small ? n32 : lb1 + x1.m;
  const bool ok_r = all_small ? x1_last : run_ends;
  // right, looking for the end of the run (x1.c == 0 whenever x1.m > 0: the entries are sorted)
  const uint32_t ub_e = 4u * x1.s + x1.m;
  const bool ok_e = x1.m < v1 || x1_last;
  // left
  const bool further_left = x1.c == 0u && x1.s != 0u;
  const bool into_s0 = x1.c + x1.m == 4u;  // the matches (if any) reach the end of s1 and go on in s0
  const uint32_t ub_l = into_s0 ? 4u * x0.s + x0.m : lb1 + x1.m;
  const bool ok_l = !further_left && (!into_s0 || x0.m < v0 || x0.s == last_s);
  const uint32_t lb_keep = b->lb;
  b->lb = right ? (want_lb ? lb_r : lb_keep) : lb1;
  b->ub = right ? (want_lb ? ub_r : ub_e) : ub_l;
  const bool ok = right ? (want_lb ? ok_r : ok_e) : ok_l;
  return ok ? 0 : 2;
}

// rev[predicted] when it matches the query (:164), from the classification of its own sector
__device__ __forceinline__ bool direct_match(uint32_t pred, const Sector& x, const uint32_t pos[4], uint32_t* idx) {
  const uint32_t j0 = pred & 3u;
  if (j0 < x.c || j0 >= x.c + x.m) return false;
  const uint32_t a = (j0 & 1u) ? pos[1] : pos[0], c = (j0 & 1u) ? pos[3] : pos[2];
  *idx = (j0 & 2u) ? c : a;
  return true;
}

__device__ __forceinline__ uint32_t kmer_uhadd(uint32_t a, uint32_t b) {  // floor((a + b) / 2) without overflow
#ifdef SB_HOST_SIM
  return (uint32_t)(((uint64_t)a + b) >> 1);
#else
  return __uhadd(a, b);
#endif
}
__device__ __forceinline__ uint32_t kmer_add_clamped(uint32_t a, uint32_t b, uint32_t top) {  // min(a + b, top), a <= top
  const uint32_t t = a + b;
  return (t < a || t > top) ? top : t;
}
__device__ __forceinline__ int kmer_flog2(uint32_t v) {  // floor(log2 v), v >= 1
#ifdef SB_HOST_SIM
  return 31 - __builtin_clz(v);
#else
  return 31 - __clz((int)v);
#endif
}

// Phase 2 in pieces, so that a kernel can run the one loop in it with a warp-uniform trip count (query.cu) and meet all its
// lanes again before rev[rank] is read; replay_plquery below strings the same pieces together for a single query.
constexpr uint32_t kNoRank = 0xFFFFFFFFu;  // "-1" (ranks are < n <= 2^32 - 16)
struct ReplayState {
  uint32_t rank;    // the rank whose rev[] plQuery returns, or kNoRank; final once !searching()
  uint32_t lo, hi;  // binarySearch's interval (:133); the reference only ever calls it with lo <= hi, so lo > hi = "over"
  __device__ __forceinline__ bool searching() const { return lo <= hi; }
};

// One level of binarySearch (:133-153) over {lb, ub}, written as selects; a no-op once the search is over, so that the
// lanes of a warp can step together.  hi == lo + 2 returns lo + 1 unverified (:136), which is the mid of that interval:
// "base case or match at mid" is one outcome, rank = mid.
__device__ __forceinline__ void replay_search_step(const Bounds& b, ReplayState* rs) {
  const uint32_t lo = rs->lo, hi = rs->hi;
  const uint32_t mid = kmer_uhadd(lo, hi);
  // (bitwise on purpose: no short-circuit branches inside the warp-uniform loop)
  const bool found = (lo <= hi) & ((hi - lo == 2u) | ((mid >= b.lb) & (mid < b.ub)));  // :136, :141
  const bool over = found | (lo + 1u >= hi);                                           // :142 (true as well once lo > hi)
  const bool small = mid < b.lb;                                                    // :143-147 / :148-152
  rs->rank = found ? mid : rs->rank;
  rs->lo = over ? 1u : (small ? mid : lo);
  rs->hi = over ? 0u : (small ? hi : mid);
}

// A long search whose steps all go right (every mid below lb: the search from rank 0 that the (int)predicted arithmetic
// of :209,:225 causes for predicted ranks >= 2^31) is jumped in closed form: after j such steps
// lo = hi - ceil(D / 2^j), D = hi - lo.  Those mids are below lb (no match, "too small") as long as ceil(D / 2^j) >
// hi - lb, the base case hi == lo + 2 (:136) and the empty-interval exit (:142) need ceil(D / 2^(j-1)) >= 3; both hold
// for ceil(D / 2^j) >= G = max(hi - lb + 1, 2), and j = floor(log2(D - 1)) - floor(log2(G - 1)) - 1 guarantees that.
__device__ __forceinline__ void replay_search_jump(const Bounds& b, ReplayState* rs) {
  const uint32_t lo = rs->lo, hi = rs->hi;
  if (hi > lo && hi - lo > 32u) {
    const uint32_t D = hi - lo;
    const uint32_t T = b.lb > hi ? 0u : hi - b.lb;
    const uint32_t G = T + 1u > 2u ? T + 1u : 2u;
    const int j = kmer_flog2(D - 1u) - kmer_flog2(G - 1u) - 1;
    if (j >= 1) rs->lo = hi - (((D - 1u) >> j) + 1u);
  }
}

// plQuery (:159-248) for s.length() == length == k (no gallop loops) over {lb, ub}, up to the call of binarySearch (:245).
// pred < n.  Both sides' windows are computed for every lane and selected (the lanes of a warp go left and right in equal
// numbers).
__device__ __forceinline__ void replay_windows(const IndexView& ix, uint32_t pred, const Bounds& b, ReplayState* rs) {
  const uint32_t nm1 = (uint32_t)ix.n - 1u;
  auto is_match = [&](uint32_t r) { return r >= b.lb && r < b.ub; };
  // look right (:167-204): hi = min(n-1, predicted + mostOver), then min(n-1, predicted + maxOver + 1)
  const uint32_t hi1 = kmer_add_clamped(pred, (uint32_t)ix.mostOver, nm1);
  const uint32_t hi2 = kmer_add_clamped(pred, (uint32_t)ix.maxOver + 1u, nm1);
  // look left (:206-243): lo = max(0, (int)predicted - mostUnder), then max(0, (int)predicted - maxUnder - 1); the (int)
  // cast wraps negative for predicted >= 2^31, which compat mode keeps
  uint32_t lo1, lo2;
  if (ix.compat) {
    const int32_t v1 = (int32_t)(pred - (uint32_t)ix.mostUnder);
    const int32_t v2 = (int32_t)(pred - (uint32_t)ix.maxUnder - 1u);
    lo1 = (uint32_t)(v1 > 0 ? v1 : 0);
    lo2 = (uint32_t)(v2 > 0 ? v2 : 0);
  } else {
    const uint32_t d1 = (uint32_t)ix.mostUnder, d2 = (uint32_t)ix.maxUnder + 1u;
    lo1 = pred > d1 ? pred - d1 : 0u;
    lo2 = pred > d2 ? pred - d2 : 0u;
  }
  const bool right = pred < b.lb;          // suffix smaller than the query (:167)
  const uint32_t p1 = right ? hi1 : lo1;   // the first bound probed (:171 / :209)
  const uint32_t p2 = right ? hi2 : lo2;   // the second one (:180 / :225)
  // the second bound is probed when the first is still on the same side of the query as the prediction (:175 / :220)
  const bool second = right ? (p1 < b.lb) : !(p1 < b.lb);
  const bool m0 = is_match(pred), m1 = is_match(p1), m2 = second && is_match(p2);  // :164, :174 / :213, :183 / :228
  const uint32_t lo = right ? (second ? hi1 : pred) : (second ? lo2 : lo1);
  const uint32_t hi = right ? (second ? hi2 : hi1) : (second ? lo1 : pred);
  // lo > hi happens in one way only: the (int) wrap of :209 sends lo1 to 0 while :225 does not wrap, for a query smaller
  // than every suffix.  binarySearch (size_t arithmetic) then probes (lo + hi) / 2 once and gives up (:141-142).
  const uint32_t mid = kmer_uhadd(lo, hi);
  const bool m3 = lo > hi && is_match(mid);
  const bool decided = m0 || m1 || m2 || lo > hi;
  rs->rank = m0 ? pred : (m1 ? p1 : (m2 ? p2 : (m3 ? mid : kNoRank)));
  rs->lo = decided ? 1u : lo;
  rs->hi = decided ? 0u : hi;
}

// Phase 2 for one query: the rank whose rev[] plQuery returns, or -1.
__device__ __forceinline__ long long replay_plquery(const IndexView& ix, uint32_t pred, const Bounds& b) {
  ReplayState rs;
  replay_windows(ix, pred, b, &rs);
  replay_search_jump(b, &rs);
  while (rs.searching()) replay_search_step(b, &rs);  // :245
  return rs.rank == kNoRank ? -1ll : (long long)rs.rank;
}

// rev[r] (the reference's suffix array, sapling_api.h:41) read from the rank lines
__device__ __forceinline__ uint32_t rev_at(const IndexView& ix, uint64_t r, uint64_t pol) {
  return ld_u32_pol(ix.lines + (r >> 2) * 8u + 4u + (r & 3u), pol);
}

// The answer once the bounds are known: phase 2, then rev[rank] -- in a line this lane has just read.
__device__ __forceinline__ long long finish_kmer(const IndexView& ix, uint32_t pred, const Bounds& b, const L2Policies& pol) {
  ReplayState rs;
  replay_windows(ix, pred, b, &rs);
  replay_search_jump(b, &rs);
  while (rs.searching()) replay_search_step(b, &rs);  // :245
  uint32_t rank = rs.rank;
  SB_MEET(rank);  // the lanes leave the loop after different numbers of steps: one read for all of them
  if (rank == kNoRank) return -1;  // :246
  SB_SIM_ADD(g_sim_final_loads, 1);
  return (long long)rev_at(ix, rank, pol.sa);  // :247
}
// The whole path for one k-mer x whose predicted rank is pred (< n): plQuery's return value.  (The general search on
// its own; the kernels try the two-sector shortcut first and come here for what it leaves.)
template <bool kTies>
__device__ __forceinline__ long long answer_kmer(const IndexView& ix, uint64_t x, uint32_t pred, const L2Policies& pol) {
  const KmerKey key = make_key<kTies>(ix, x);
  Search se;
  se.begin(ix, pred);
  Bounds b;
  b.lb = b.ub = 0;
  bool is_first = true, direct = false;
  uint32_t idx = 0;
  // one way out of the loop: the lanes of a warp leave it after different numbers of sectors and meet again behind it,
  // before phase 2, instead of each group running phase 2 on its own
  for (;;) {
    uint32_t pos[4];
    const Sector sc = classify_sector<kTies>(ix, key, se.t, pol, pos);
    direct = is_first && direct_match(pred, sc, pos, &idx);  // :164: a third of all queries
    if (direct || se.feed(ix, pred, sc, is_first, &b)) break;
    is_first = false;
  }
  SB_MEET(idx);
  if (direct) return (long long)idx;
  return finish_kmer(ix, pred, b, pol);
}

// The same answer by the kernels' schedule: sector of the prediction, one neighbour, and only then the general search.
template <bool kTies>
__device__ __forceinline__ long long answer_kmer_fast(const IndexView& ix, uint64_t x, uint32_t pred, const L2Policies& pol) {
  const KmerKey key = make_key<kTies>(ix, x);
  uint32_t pos[4], idx, neighbour;
  const Sector s0 = classify_sector<kTies>(ix, key, pred >> 2, pol, pos);
  if (direct_match(pred, s0, pos, &idx)) return (long long)idx;  // :164
  Bounds b;
  int st = two_sector_first(ix, s0, &b, &neighbour);
  if (st == 1) {
    const Sector s1 = classify_sector<kTies>(ix, key, neighbour, pol, pos);
    st = two_sector_second(ix, s0, s1, &b);
  }
  if (st == 0) return finish_kmer(ix, pred, b, pol);
  return answer_kmer<kTies>(ix, x, pred, pol);
}

}  // namespace sb

// sim_query.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the *device* query code (sapling_b200/csrc/kmer.cuh, query.cuh, common.cuh) for the host with g++
// by supplying host stand-ins for the handful of CUDA intrinsics it uses, so that the query path can be
// differential-tested against the oracle on a machine without a GPU.  It is never linked into libsapling_b200.so
// and never used by the product.
#define SB_HOST_SIM 1
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline double __ll2double_rn(long long x) { return (double)x; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline long long __double2ll_rz(double a) { return (long long)a; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; *p += v; return o;
}

#include "../../sapling_b200/csrc/kmer.cuh"
#include "../../sapling_b200/csrc/query.cuh"

namespace sb { void set_error(const char*, ...) {} const char* last_error() { return ""; } }

namespace {
// rank lines built by the same pack_rank_sector the build kernel runs
uint32_t* build_lines(const uint64_t* genome, const uint32_t* sa, uint64_t n, int bases, unsigned long long* n_escapes) {
  const uint64_t sectors = sb::line_sectors(n);
  uint32_t* lines = new uint32_t[sectors * 8];
  unsigned long long esc = 0;
  for (uint64_t s = 0; s < sectors; s++) {
    sb::pack_rank_sector(genome, sa, n, bases, s * 4, lines + s * 8);
    const uint64_t D = ((uint64_t)lines[s * 8 + 3] << 32) | lines[s * 8 + 2];
    if (s * 4 < n) {
      esc += (D >> 63);
      for (int j = 1; j < 4; j++) esc += ((D >> (21 * (j - 1))) & 0x1FFFFFu) == 0x1FFFFFu && s * 4 + j < n;
    }
  }
  if (n_escapes) *n_escapes = esc;
  return lines;
}
sb::IndexView make_view(const uint64_t* genome, const uint32_t* lines, int bases, const int64_t* model_xy, uint64_t n, int k,
                        int nb, const int* five, int compat, unsigned long long* oob, const uint2* narrow,
                        const int64_t* last_xy) {
  sb::IndexView ix;
  ix.genome = genome; ix.lines = lines; ix.line_bases = bases;
  ix.model = reinterpret_cast<const sb::ModelEntry*>(model_xy);
  ix.narrow = narrow; ix.last_x = last_xy ? last_xy[0] : 0; ix.last_y = last_xy ? last_xy[1] : 0;
  ix.n = n; ix.k = k; ix.nb = nb; ix.shift = 2 * k - nb;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat; ix.oob_counter = oob; ix.hints = 0;
  return ix;
}
template <bool kTies>
long long staged_answer(const sb::IndexView& ix, uint64_t x, uint32_t pred, const sb::L2Policies& pol) {
  const sb::KmerKey key = sb::make_key<kTies>(ix, x);
  sb::Search se;
  se.begin(ix, pred);
  sb::Bounds b;
  uint32_t pos[4], idx;
  sb::Sector sc = sb::classify_sector<kTies>(ix, key, se.t, pol, pos);
  if (sb::direct_match(pred, sc, pos, &idx)) return (long long)idx;
  bool resolved = se.feed(ix, pred, sc, true, &b);
  if (!resolved) {
    sc = sb::classify_sector<kTies>(ix, key, se.t, pol, pos);
    resolved = se.feed(ix, pred, sc, false, &b);
  }
  while (!resolved) {
    const uint4 w = se.pack();
    const uint32_t t = se.t;
    sb::Search s2;
    s2.begin(ix, 0);
    s2.unpack(w, t);
    sc = sb::classify_sector<kTies>(ix, key, s2.t, pol, pos);
    resolved = s2.feed(ix, pred, sc, false, &b);
    se = s2;
  }
  return sb::finish_kmer(ix, pred, b, pol);
}

}  // namespace

extern "C" {

void sim_skip_counters(unsigned long long* tried, unsigned long long* ok) {
  *tried = sb::g_sim_skip_tried;
  *ok = sb::g_sim_skip_ok;
}
// sector loads / genome-decided entries / extra rev[] loads of the k-mer path so far
void sim_kmer_counters(unsigned long long* sectors, unsigned long long* slow, unsigned long long* finals) {
  *sectors = sb::g_sim_sector_loads;
  *slow = sb::g_sim_slow_entries;
  *finals = sb::g_sim_final_loads;
}

// The batch k-mer path (kmer.cuh answer_kmer) exactly as the kernels call it.
// genome: packed words (with pad), sa: n entries, model: interleaved {x,y} x ((1<<nb)+1); bases: leading bases per
// rank-line entry (0 = the library's choice for n).  out_lines may be NULL; otherwise it receives line_sectors(n) * 8
// words (to compare with the GPU's).
void sim_kmer_answer(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                     const int* five, int compat, const uint64_t* kmers, size_t nq, int64_t* out,
                     unsigned long long* oob, const uint2* narrow, const int64_t* last_xy, int bases,
                     uint32_t* out_lines, unsigned long long* n_escapes) {
  if (bases <= 0) bases = sb::line_bases_for(n);
  uint32_t* lines = build_lines(genome, sa, n, bases, n_escapes);
  if (out_lines) memcpy(out_lines, lines, sb::line_sectors(n) * 32);
  const sb::IndexView ix = make_view(genome, lines, bases, model_xy, n, k, nb, five, compat, oob, narrow, last_xy);
  const sb::L2Policies pol = sb::make_policies(0);
  const uint64_t kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
  for (size_t i = 0; i < nq; i++) {
    const uint64_t x = kmers[i] & kmask;
    const uint32_t pred = (uint32_t)sb::clamp_prediction(ix, sb::predict_rank(ix, x, pol.model));
    out[i] = k > bases ? sb::answer_kmer<true>(ix, x, pred, pol) : sb::answer_kmer<false>(ix, x, pred, pol);
    // the partitioned kernel's schedule of the same search (query.cu kmer_query_ordered_kernel): two rounds in place,
    // then the state travels through its packed form (the warp's shared-memory stack) between rounds
    const long long staged = k > bases ? staged_answer<true>(ix, x, pred, pol) : staged_answer<false>(ix, x, pred, pol);
    if (staged != out[i]) out[i] = -12345;  // poison: the caller's comparison with the oracle then fails
    // and the two-sector shortcut in front of it
    const long long fast = k > bases ? sb::answer_kmer_fast<true>(ix, x, pred, pol) : sb::answer_kmer_fast<false>(ix, x, pred, pol);
    if (fast != out[i]) out[i] = -12346;
  }
  delete[] lines;
}

// The literal replay (query.cuh Replay) for k-mers: what the string kernel and the probe counter run.
// probes != NULL: receives the total number of getLcp calls (kSkip off), as bench.py's device counter does.
void sim_kmer_batch(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                    const int* five, int compat, const uint64_t* kmers, size_t nq, int64_t* out,
                    unsigned long long* oob, const uint2* narrow, const int64_t* last_xy, unsigned long long* probes) {
  const int bases = sb::line_bases_for(n);
  uint32_t* lines = build_lines(genome, sa, n, bases, nullptr);
  const sb::IndexView ix = make_view(genome, lines, bases, model_xy, n, k, nb, five, compat, oob, narrow, last_xy);
  const sb::L2Policies pol = sb::make_policies(0);
  unsigned long long np = 0;
  for (size_t i = 0; i < nq; i++) {
    sb::KmerQuery q; q.q = kmers[i] << (64 - 2 * k); q.k = (uint32_t)k;
    if (!probes) {
      out[i] = sb::pl_query<false>(ix, q, kmers[i]);
    } else {
      uint64_t pred = sb::predict_rank(ix, kmers[i], pol.model);
      if (pred >= n) pred = n - 1;
      sb::Replay<false, sb::KmerQuery, false> rp;
      rp.begin(pred);
      long long result;
      for (;;) {
        const bool final_read = rp.state == sb::ST_FINAL;
        const bool done = rp.step(ix, q, pol, &result);
        if (!final_read) np++;
        if (done) break;
      }
      out[i] = result;
    }
  }
  if (probes) *probes = np;
  delete[] lines;
}

void sim_string_batch(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                      const int* five, int compat, const uint64_t* words, const uint64_t* word_off,
                      const uint32_t* slens, const uint32_t* lengths, const int64_t* kmers, size_t nq, int64_t* out,
                      unsigned long long* oob, const uint2* narrow, const int64_t* last_xy) {
  const int bases = sb::line_bases_for(n);
  uint32_t* lines = build_lines(genome, sa, n, bases, nullptr);
  const sb::IndexView ix = make_view(genome, lines, bases, model_xy, n, k, nb, five, compat, oob, narrow, last_xy);
  for (size_t i = 0; i < nq; i++) {
    sb::StringQuery q; q.w = words + word_off[i]; q.slen_ = slens[i]; q.length_ = lengths[i];
    out[i] = sb::pl_query<true>(ix, q, (uint64_t)kmers[i]);
  }
  delete[] lines;
}

// Phase 2 of the k-mer path on its own (kmer.cuh replay_plquery: the reference's control flow over {lb, ub}) against a
// direct transcription of the reference's plQuery / binarySearch (sapling_api.h:133-248) with the same abstract
// comparator -- size_t arithmetic, the (int) casts of :209 and :225, recursion unrolled, nothing jumped.  No genome is
// needed, so ranks >= 2^31 and windows of any size are testable on the CPU.  Returns the number of cases in
// which the two disagree; *first_bad receives the index of the first one.
static long long literal_plquery(uint64_t n, uint64_t pred, uint64_t lb, uint64_t ub, const int* five, int compat) {
  const long long maxOver = five[0], maxUnder = five[1], mostOver = five[3], mostUnder = five[4];
  auto small = [&](uint64_t r) { return r < lb; };
  auto match = [&](uint64_t r) { return r >= lb && r < ub; };
  if (match(pred)) return (long long)pred;                                   // :164
  uint64_t lo, hi;
  if (small(pred)) {                                                          // :167
    lo = pred;
    hi = pred + (uint64_t)mostOver < n - 1 ? pred + (uint64_t)mostOver : n - 1;  // :171
    if (match(hi)) return (long long)hi;                                      // :174
    if (small(hi)) {                                                          // :175
      lo = hi;
      hi = pred + (uint64_t)maxOver + 1 < n - 1 ? pred + (uint64_t)maxOver + 1 : n - 1;  // :180
      if (match(hi)) return (long long)hi;                                    // :183
    }
  } else {
    if (compat) {
      const int v = (int)(unsigned)pred - (int)mostUnder;                     // :209  (int)predicted-mostUnder
      lo = (uint64_t)(v > 0 ? v : 0);
    } else {
      lo = pred > (uint64_t)mostUnder ? pred - (uint64_t)mostUnder : 0;
    }
    hi = pred;
    if (match(lo)) return (long long)lo;                                      // :213
    if (!small(lo)) {                                                         // :220
      hi = lo;
      if (compat) {
        const int v = (int)((unsigned)pred - (unsigned)maxUnder - 1u);        // :225 (two's complement wrap, as compiled)
        lo = (uint64_t)(v > 0 ? v : 0);
      } else {
        lo = pred > (uint64_t)maxUnder + 1 ? pred - (uint64_t)maxUnder - 1 : 0;
      }
      if (match(lo)) return (long long)lo;                                    // :228
    }
  }
  for (;;) {                                                                  // binarySearch :133-153
    if (hi == lo + 2) return (long long)(lo + 1);
    const uint64_t mid = (lo + hi) >> 1;
    if (match(mid)) return (long long)mid;
    if (lo + 1 >= hi) return -1;
    if (small(mid)) lo = mid; else hi = mid;
  }
}
uint64_t sim_replay_abstract(uint64_t n, const uint64_t* pred, const uint64_t* lb, const uint64_t* ub, size_t count,
                             const int* five, int compat, uint64_t* first_bad) {
  sb::IndexView ix;
  memset(&ix, 0, sizeof(ix));
  ix.n = n; ix.k = 21;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat;
  uint64_t bad = 0;
  for (size_t i = 0; i < count; i++) {
    sb::Bounds b;
    b.lb = (uint32_t)lb[i];
    b.ub = (uint32_t)ub[i];
    const long long got = sb::replay_plquery(ix, (uint32_t)pred[i], b);
    const long long exp = literal_plquery(n, pred[i], lb[i], ub[i], five, compat);
    if (got != exp && bad++ == 0 && first_bad) *first_bad = i;
  }
  return bad;
}
}

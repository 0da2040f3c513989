// sim_query.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the *device* query code (sapling_b200/csrc/query.cuh, common.cuh) for the host with g++
// by supplying host stand-ins for the handful of CUDA intrinsics it uses, so that the control-flow
// replay can be differential-tested against the oracle on a machine without a GPU.  It is never
// linked into libsapling_b200.so and never used by the product.
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

#include "../../sapling_b200/csrc/query.cuh"

namespace sb { void set_error(const char*, ...) {} const char* last_error() { return ""; } }

extern "C" {

void sim_skip_counters(unsigned long long* tried, unsigned long long* ok) {
  *tried = sb::g_sim_skip_tried;
  *ok = sb::g_sim_skip_ok;
}

// genome: packed words (with pad), sa: n entries, model: interleaved {x,y} x ((1<<nb)+1)
void sim_kmer_batch(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                    const int* five, int compat, const uint64_t* kmers, size_t nq, int64_t* out,
                    unsigned long long* oob, const uint2* narrow, const int64_t* last_xy) {
  sb::IndexView ix;
  ix.genome = genome; ix.sa = sa; ix.model = reinterpret_cast<const sb::ModelEntry*>(model_xy);
  ix.n = n; ix.k = k; ix.nb = nb; ix.shift = 2 * k - nb;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat; ix.oob_counter = oob;
  ix.narrow = narrow; ix.last_x = last_xy[0]; ix.last_y = last_xy[1]; ix.hints = 0;
  for (size_t i = 0; i < nq; i++) {
    sb::KmerQuery q; q.q = kmers[i] << (64 - 2 * k); q.k = (uint32_t)k;
    out[i] = sb::pl_query<false>(ix, q, kmers[i]);
  }
}

// rank lines (common.cuh IndexView) built by the same pack_rank_sector the build kernel runs, then the packed-mode
// replay.  out_packed may be NULL; otherwise it receives packed_sectors(n, shift) * 8 words (to compare with the GPU's).
void sim_kmer_batch_packed(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                           const int* five, int compat, const uint64_t* kmers, size_t nq, int64_t* out,
                           unsigned long long* oob, int bases, int shift, uint32_t* out_packed,
                           unsigned long long* n_escapes) {
  const uint64_t sectors = sb::packed_sectors(n, shift);
  uint32_t* packed = out_packed ? out_packed : new uint32_t[sectors * 8];
  unsigned long long esc = 0;
  for (uint64_t s = 0; s < sectors; s++) {
    const uint64_t r0 = ((s >> 2) << shift) + ((s & 3u) << 2);
    sb::pack_rank_sector(genome, sa, n, bases, r0, packed + s * 8);
    const uint64_t D = ((uint64_t)packed[s * 8 + 3] << 32) | packed[s * 8 + 2];
    if (r0 < n) {
      esc += (D >> 63);
      for (int j = 1; j < 4; j++) esc += ((D >> (21 * (j - 1))) & 0x1FFFFFu) == 0x1FFFFFu && r0 + j < n;
    }
  }
  if (n_escapes) *n_escapes = esc;
  sb::IndexView ix;
  ix.genome = genome; ix.sa = sa; ix.ext = nullptr; ix.ext_bases = 0;
  ix.packed = packed; ix.packed_bases = bases; ix.packed_shift = shift;
  ix.model = reinterpret_cast<const sb::ModelEntry*>(model_xy);
  ix.n = n; ix.k = k; ix.nb = nb; ix.shift = 2 * k - nb;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat; ix.oob_counter = oob;
  ix.narrow = nullptr; ix.last_x = 0; ix.last_y = 0; ix.hints = 0;
  const sb::L2Policies pol = sb::make_policies(0);
  for (size_t i = 0; i < nq; i++) {
    sb::KmerQuery q; q.q = kmers[i] << (64 - 2 * k); q.k = (uint32_t)k;
    const uint64_t pred = sb::clamp_prediction(ix, sb::predict_rank(ix, kmers[i], pol.model));
    sb::SaPacked sp;
    sp.anchor(ix, pred);
    out[i] = sb::pl_query_from<false, false, sb::KmerQuery, sb::SaPacked, true, 2>(ix, q, pred, 0, pol, sp);
  }
  if (!out_packed) delete[] packed;
}

// The lean 32-bit k-mer replay (query.cuh kmer_replay32) on the three layouts.  mode 0: suffix-array sector + packed
// genome; 1: inline-prefix entries (ext_bases leading bases, k <= ext_bases); 2: rank lines (bases, shift).
int sim_kmer_batch_lean(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                        const int* five, int compat, const uint64_t* kmers, size_t nq, int64_t* out,
                        unsigned long long* oob, int mode, int bases, int shift) {
  sb::IndexView ix;
  ix.genome = genome; ix.sa = sa; ix.ext = nullptr; ix.ext_bases = 0;
  ix.packed = nullptr; ix.packed_bases = bases; ix.packed_shift = shift;
  ix.model = reinterpret_cast<const sb::ModelEntry*>(model_xy);
  ix.n = n; ix.k = k; ix.nb = nb; ix.shift = 2 * k - nb;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat; ix.oob_counter = oob;
  ix.narrow = nullptr; ix.last_x = 0; ix.last_y = 0; ix.hints = 0;
  if (!sb::lean_eligible(ix)) return -1;
  uint32_t* packed = nullptr;
  sb::ExtEntry* ext = nullptr;
  // the device arrays are padded to whole lines; sectors are read whole
  uint32_t* sa_pad = new uint32_t[sb::sa_alloc_entries(n) + 16]();
  memcpy(sa_pad, sa, n * sizeof(uint32_t));
  ix.sa = sa_pad;
  if (mode == 4 && shift != 4) return -2;  // the flat replay reads tiling lines only
  if (mode == 2 || mode == 3 || mode == 4 || mode == 5) {
    const uint64_t sectors = sb::packed_sectors(n, shift);
    packed = new uint32_t[sectors * 8];
    for (uint64_t s = 0; s < sectors; s++) {
      const uint64_t r0 = ((s >> 2) << shift) + ((s & 3u) << 2);
      sb::pack_rank_sector(genome, sa, n, bases, r0, packed + s * 8);
    }
    ix.packed = packed;
  } else if (mode == 1) {
    ext = new sb::ExtEntry[n];
    for (uint64_t r = 0; r < n; r++) {
      ext[r].pos = sa[r];
      ext[r].reserved = 0;
      const uint64_t w = sb::load_bases32(genome, sa[r]);
      ext[r].prefix = bases >= 32 ? w : (w >> (64 - 2 * bases)) << (64 - 2 * bases);
    }
    ix.ext = ext;
    ix.ext_bases = bases;
  }
  const sb::L2Policies pol = sb::make_policies(0);
  for (size_t i = 0; i < nq; i++) {
    const uint64_t q = kmers[i] << (64 - 2 * k);
    const uint32_t pred = (uint32_t)sb::clamp_prediction(ix, sb::predict_rank(ix, kmers[i], pol.model));
    if (mode == 4) {
      out[i] = sb::kmer_replay_flat<true>(ix, q, pred, pol);
    } else if (mode == 5) {  // the replay cut in two (parked tails): three probes, then binarySearch resumed cold
      sb::SaPacked32 sp;
      sp.anchor(ix, pred);
      sb::Lean32 st;
      long long r = 0;
      if (!sb::kmer_replay32_head<2, true>(ix, q, pred, pol, sp, st, &r)) {
        sb::SaPacked32 cold;
        cold.abase = 0xFFFFFFF0u;
        cold.cur = 0xFFFFFFFFu;
        r = sb::kmer_replay32_tail<2, true>(ix, q, pol, cold, st);
      }
      out[i] = r;
    } else if (mode == 2) {
      sb::SaPacked32 sp;
      sp.anchor(ix, pred);
      out[i] = sb::kmer_replay32<2, true>(ix, q, pred, pol, sp);
    } else if (mode == 3) {  // anchor line staged in the thread's shared-memory slot
      uint4 slot[sb::kLineSlotU4];
      sb::SaLine32 sl;
      sl.sm = slot;
      sl.anchor(ix, pred, pol.sa);
      out[i] = sb::kmer_replay32<2, true>(ix, q, pred, pol, sl);
    } else if (mode == 1) {
      sb::SaNone32 none;
      out[i] = sb::kmer_replay32<1, true>(ix, q, pred, pol, none);
    } else {
      sb::SaSector32 ss;
      ss.fill(ix, pred, pol.sa);
      out[i] = sb::kmer_replay32<0, true>(ix, q, pred, pol, ss);
    }
  }
  delete[] packed;
  delete[] ext;
  delete[] sa_pad;
  return 0;
}

void sim_string_batch(const uint64_t* genome, const uint32_t* sa, const int64_t* model_xy, uint64_t n, int k, int nb,
                      const int* five, int compat, const uint64_t* words, const uint64_t* word_off,
                      const uint32_t* slens, const uint32_t* lengths, const int64_t* kmers, size_t nq, int64_t* out,
                      unsigned long long* oob, const uint2* narrow, const int64_t* last_xy) {
  sb::IndexView ix;
  ix.genome = genome; ix.sa = sa; ix.model = reinterpret_cast<const sb::ModelEntry*>(model_xy);
  ix.n = n; ix.k = k; ix.nb = nb; ix.shift = 2 * k - nb;
  ix.maxOver = five[0]; ix.maxUnder = five[1]; ix.mostOver = five[3]; ix.mostUnder = five[4];
  ix.compat = compat; ix.oob_counter = oob;
  ix.narrow = narrow; ix.last_x = last_xy[0]; ix.last_y = last_xy[1]; ix.hints = 0;
  for (size_t i = 0; i < nq; i++) {
    sb::StringQuery q; q.w = words + word_off[i]; q.slen_ = slens[i]; q.length_ = lengths[i];
    out[i] = sb::pl_query<true>(ix, q, (uint64_t)kmers[i]);
  }
}
}

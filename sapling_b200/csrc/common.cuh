// common.cuh -- shared definitions for the sapling_b200 CUDA library (sm_100a).
//
// Data layout in HBM:
//   genome : 2-bit packed, 32 bases per 64-bit word, base i in bits [63-2(i%32)-1, 63-2(i%32)]
//            of word i/32 ("big-endian in word": integer order of a word == lexicographic order
//            of its 32 bases).  A=0 C=1 G=2 T=3 (the vals[] table of the reference,
//            sapling_api.h:494-498).  Padded with GENOME_PAD_WORDS zero words.
//   lines  : rank lines -- the reference's `rev` (sapling_api.h:41) together with the leading bases of every suffix,
//            16 ranks per 128-byte line (below).  The ONLY resident copy of the suffix array.
//   narrow : uint2[1<<nb] = {xoff, y}: the checkpoints xlist/ylist (sapling_api.h:65) in 8 bytes per bucket (model.cu);
//            model : ModelEntry[(1<<nb)+1] = {x, y}, resident only when the narrow form cannot represent the table.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int GENOME_PAD_WORDS = 4;

struct __align__(16) ModelEntry {
  long long x;
  long long y;
};

// Rank lines: the layout built around the one fact that decides this path on B200 -- an L2 miss fills a whole 128-byte
// DRAM line whatever the width of the load (profiles/r1_gather_dram_granularity.txt), so the cost of a query is the number
// of distinct LINES it touches.  A rank line holds 16 consecutive ranks as four self-contained 32-byte sectors; sector s
// holds ranks 4s .. 4s+3 and classifies all four against a query on its own (kmer.cuh):
//   v[0..1]  P0  = the first `line_bases` bases of the suffix at the sector's first rank, as a 2*line_bases-bit integer
//   v[2..3]  D   = three 21-bit deltas d1,d2,d3 (bits 0-20, 21-41, 42-62): P_j = P0 + d_j   (suffixes are sorted, so d_j >= 0)
//                  d_j == kPackedEscape, or bit 63 for entry 0: "compare this entry against the packed genome instead"
//                  (delta overflow, a suffix that ends within 32 bases of the end of the text, padding past the last rank)
//   v[4..7]  the four text positions (the reference's rev[r])
constexpr uint32_t kPackedEscape = 0x1FFFFFu;
constexpr int kPackedDeltaBits = 21;
constexpr int kLineMaxBases = 31;  // 2 * bases <= 62 bits: differences of prefixes fit a signed 64-bit integer

// Everything the query kernels read; passed by value as a kernel argument (lives in the constant bank, so the scalars
// cost no memory traffic).
struct IndexView {
  const uint64_t* genome;
  const uint32_t* lines;   // rank lines (see above)
  int line_bases;          // leading bases per entry (<= kLineMaxBases)
  const ModelEntry* model; // wide checkpoints; nullptr when `narrow` represents the table
  const uint2* narrow;     // 8 bytes per bucket (model.cu); nullptr -> use the wide `model` table
  long long last_x, last_y;  // checkpoint (1<<nb): the largest k-mer of the genome
  uint64_t n;
  int k;
  int nb;
  int shift;  // 2k - nb
  int maxOver, maxUnder, mostOver, mostUnder;
  int compat;  // 1: the reference's (int)predicted window arithmetic (sapling_api.h:209,225)
  unsigned long long* oob_counter;  // incremented when predicted >= n (reference: UB)
  unsigned hints;  // L2 residency hints (HINT_* bits)
};

enum : unsigned {
  HINT_GENOME_KEEP = 1u,   // packed genome loads: L2 evict_last
  HINT_MODEL_KEEP = 2u,    // model loads: L2 evict_last
  HINT_SA_STREAM = 4u,     // rank-line loads: L2 evict_first
  HINT_IO_STREAM = 8u,     // k-mer reads / result writes: ld.cs / st.cs
  HINT_SA_KEEP = 16u       // rank-line loads: L2 evict_last (wins over HINT_SA_STREAM)
};

#define SB_CUDA_CHECK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      sb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                    cudaGetErrorString(_e));                                         \
      return -1;                                                                     \
    }                                                                                \
  } while (0)

void set_error(const char* fmt, ...);
const char* last_error();

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// L2 eviction-priority policies for ld.global.nc.L2::cache_hint (PTX createpolicy).  The index is
// far larger than L2 and is gathered at 32-byte granularity, so without hints the random SA / model
// sectors flush the small, hot packed genome out of L2 (ncu: 27% L2 hit rate, profiles/).
#ifdef SB_HOST_SIM
// tests/sim: the same code compiled for the host (no PTX); policies are inert
struct L2Policies {
  uint64_t genome, model, sa;
};
inline L2Policies make_policies(unsigned) { return L2Policies{0, 0, 0}; }
inline uint64_t ld_u64_pol(const uint64_t* p, uint64_t) { return *p; }
inline uint4 ld_u32x4_pol(const uint4* p, uint64_t) { return *p; }
inline uint32_t ld_u32_pol(const uint32_t* p, uint64_t) { return *p; }
inline uint2 ld_u32x2_pol(const uint2* p, uint64_t) { return *p; }
inline longlong2 ld_s64x2_pol(const longlong2* p, uint64_t) { return *p; }
struct U32x8 {
  uint32_t v[8];
};
inline U32x8 ld_u32x8_pol(const uint32_t* p, uint64_t) {
  U32x8 r;
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}
#else
struct L2Policies {
  uint64_t genome, model, sa;
};
__device__ __forceinline__ L2Policies make_policies(unsigned hints) {
  uint64_t keep, first, normal;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(first));
  asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(normal));
  L2Policies p;
  p.genome = (hints & HINT_GENOME_KEEP) ? keep : normal;
  p.model = (hints & HINT_MODEL_KEEP) ? keep : normal;
  p.sa = (hints & HINT_SA_KEEP) ? keep : (hints & HINT_SA_STREAM) ? first : normal;
  return p;
}
__device__ __forceinline__ uint64_t ld_u64_pol(const uint64_t* p, uint64_t pol) {
  uint64_t v;
  asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ld_u32_pol(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
// one 256-bit load (LDG.E.256, sm_100+): a whole 32-byte sector in ONE request
struct U32x8 {
  uint32_t v[8];
};
__device__ __forceinline__ U32x8 ld_u32x8_pol(const uint32_t* p, uint64_t pol) {
  U32x8 r;
  asm("ld.global.nc.L2::cache_hint.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint4 ld_u32x4_pol(const uint4* p, uint64_t pol) {
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint2 ld_u32x2_pol(const uint2* p, uint64_t pol) {
  uint2 v;
  asm("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ longlong2 ld_s64x2_pol(const longlong2* p, uint64_t pol) {
  longlong2 v;
  asm("ld.global.nc.L2::cache_hint.v2.s64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
  return v;
}
#endif

// 32 bases starting at text position idx, left-aligned in a 64-bit word (base idx in the top
// two bits).  Reads genome words idx/32 and idx/32+1 (the pad makes the second read safe).
__device__ __forceinline__ uint64_t load_bases32(const uint64_t* __restrict__ genome, uint64_t idx) {
  const uint64_t w = idx >> 5;
  const unsigned o = (unsigned)(idx & 31u) * 2u;
  const uint64_t hi = __ldg(genome + w);
  if (o == 0) return hi;
  const uint64_t lo = __ldg(genome + w + 1);
  return (hi << o) | (lo >> (64u - o));
}

// One rank-line sector (see IndexView): ranks r0 .. r0+3 -> eight 32-bit words.  Shared by the build kernel
// (sa_build.cu) and by the host simulation of the query code (tests/sim).
__device__ __forceinline__ void pack_rank_sector(const uint64_t* __restrict__ genome, const uint32_t* __restrict__ sa,
                                                 uint64_t n, int bases, uint64_t r0, uint32_t out[8]) {
  uint64_t P[4];
  bool esc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t r = r0 + (uint64_t)j;
    if (r < n) {
      const uint64_t pos = __ldg(sa + r);
      out[4 + j] = (uint32_t)pos;
      P[j] = load_bases32(genome, pos) >> (64 - 2 * bases);
      esc[j] = n - pos < 32;  // the suffix ends inside the window: its compare needs the end-of-text rules
    } else {  // padding past the last rank: never read by a query
      out[4 + j] = 0;
      P[j] = j ? P[j - 1] : 0;
      esc[j] = true;
    }
  }
  uint64_t D = esc[0] ? (1ull << 63) : 0ull;
#pragma unroll
  for (int j = 1; j < 4; j++) {
    // suffixes are sorted, so P is non-decreasing except where a short suffix was zero-padded: those are escapes
    const uint64_t d = P[j] >= P[0] ? P[j] - P[0] : (uint64_t)kPackedEscape;
    const uint64_t f = (esc[j] || d >= (uint64_t)kPackedEscape) ? (uint64_t)kPackedEscape : d;
    D |= f << (kPackedDeltaBits * (j - 1));
  }
  out[0] = (uint32_t)P[0];
  out[1] = (uint32_t)(P[0] >> 32);
  out[2] = (uint32_t)D;
  out[3] = (uint32_t)(D >> 32);
}
// sectors in a rank-line array over n ranks (whole lines, one spare)
inline uint64_t line_sectors(uint64_t n) { return (((n + 15ull) >> 4) + 1ull) * 4ull; }
// how many leading bases a 21-bit delta can carry: the mean gap between consecutive distinct prefixes, 4^bases / n,
// must stay well (16x) below 2^21 or escapes stop being rare
inline int line_bases_for(uint64_t n) {
  int lg = 0;
  while ((1ull << (lg + 1)) <= n) lg++;
  int b = (lg + 17) / 2;
  return b < 8 ? 8 : (b > kLineMaxBases ? kLineMaxBases : b);
}

// Same, but only `need` (<=32) leading bases are required: skips the second word when the first
// one already holds them.
__device__ __forceinline__ uint64_t load_bases_upto(const uint64_t* __restrict__ genome, uint64_t idx,
                                                    unsigned need) {
  const uint64_t w = idx >> 5;
  const unsigned o = (unsigned)(idx & 31u) * 2u;
  const uint64_t hi = __ldg(genome + w);
  if (o + 2u * need <= 64u) return hi << o;
  const uint64_t lo = __ldg(genome + w + 1);
  return (hi << o) | (lo >> (64u - o));
}
// same two loads with an L2 eviction policy
__device__ __forceinline__ uint64_t load_bases_upto_pol(const uint64_t* __restrict__ genome, uint64_t idx,
                                                        unsigned need, uint64_t pol) {
  const uint64_t w = idx >> 5;
  const unsigned o = (unsigned)(idx & 31u) * 2u;
  const uint64_t hi = ld_u64_pol(genome + w, pol);
  if (o + 2u * need <= 64u) return hi << o;
  const uint64_t lo = ld_u64_pol(genome + w + 1, pol);
  return (hi << o) | (lo >> (64u - o));
}
__device__ __forceinline__ uint64_t load_bases32_pol(const uint64_t* __restrict__ genome, uint64_t idx, uint64_t pol) {
  const uint64_t w = idx >> 5;
  const unsigned o = (unsigned)(idx & 31u) * 2u;
  const uint64_t hi = ld_u64_pol(genome + w, pol);
  if (o == 0) return hi;
  const uint64_t lo = ld_u64_pol(genome + w + 1, pol);
  return (hi << o) | (lo >> (64u - o));
}

// queryPiecewiseLinear (sapling_api.h:98-109).  IEEE double, one rounding per operation, no FMA
// contraction, in the reference's evaluation order:
//   (long long)( (.5 + ylo) + (yhi - ylo) * ( ((x - xlo) * 1.) / (xhi - xlo) ) ), clamped at 0.
__device__ __forceinline__ uint64_t interpolate(long long x, long long xlo, long long ylo, long long xhi,
                                                long long yhi) {
  if (xlo == xhi) return (uint64_t)ylo;
  // x == xlo (the query is its bucket's checkpoint k-mer: one query in twelve at c2) makes the quotient 0 / d, which the
  // software division answers on its slow path -- a call that ~3 lanes of every warp took, 6.6 % of all instructions
  // the query kernel issued (ncu r2f, SASS page).  The result is known without dividing: frac = +0, rise = +-0,
  // sum = .5 + ylo, truncated = ylo.  The division is fed a harmless 1 instead so that all lanes stay on the fast path.
  const long long dx = x - xlo;
  const double num = __ll2double_rn(dx == 0 ? 1 : dx);
  const double frac = __ddiv_rn(num, __ll2double_rn(xhi - xlo));
  const double rise = __dmul_rn(__ll2double_rn(yhi - ylo), frac);
  const double base = __dadd_rn(0.5, __ll2double_rn(ylo));
  const double sum = __dadd_rn(base, rise);
  long long p = __double2ll_rz(sum);
  if (dx == 0) p = ylo;
  if (p < 0) p = 0;
  return (uint64_t)p;
}

// Narrow model entry (model.cu): {xoff, y}.  xoff bit 31 clear: the bucket holds a k-mer, its checkpoint
// is x = (b << shift) + xoff.  Bit 31 set: the bucket is empty and copies the checkpoint of bucket
// b - (xoff & 0x7fffffff) (the reference's forward fill, sapling_api.h:437-449).
constexpr uint32_t kNarrowFill = 0x80000000u;

// The two narrow entries of x's bucket.
struct NarrowPair {
  uint2 e0, e1;
};
__device__ __forceinline__ NarrowPair narrow_load(const IndexView& ix, uint64_t x, uint64_t pol) {
  const uint32_t b = (uint32_t)(x >> ix.shift);  // nb <= 31 (capi.cu): bucket numbers are 32-bit
  // the table has one entry more than buckets (flagged "filled"), so bucket b + 1 is always readable: two adjacent 8-byte
  // loads from one address
  const uint2* p = ix.narrow + b;
  NarrowPair r;
  r.e0 = ld_u32x2_pol(p, pol);
  r.e1 = ld_u32x2_pol(p + 1, pol);
  return r;
}
__device__ __forceinline__ uint64_t narrow_finish(const IndexView& ix, uint64_t x, const NarrowPair& p, uint64_t pol) {
  const uint64_t b = (uint32_t)(x >> ix.shift);  // nb <= 31: fits 32 bits (the 64-bit type only serves the rare paths below)
  const uint64_t B = 1ull << ix.nb;
  // Common case -- both buckets hold a k-mer and b is not the last bucket: the three differences the interpolation needs
  // fit 32 bits (the narrow layout implies shift <= 31: xoff < 2^31, ranks < 2^32), so they are formed from the entries'
  // offsets without rebuilding the 64-bit checkpoints.  Same real values, hence the same doubles as interpolate() below.
  // (for the last bucket the "next" entry is the table's pad, whose flag is set: it takes the general path below)
  if (!((p.e0.x | p.e1.x) & kNarrowFill)) {
    const uint32_t xl = (uint32_t)x & ((1u << ix.shift) - 1u);          // x - (b << shift)
    const int32_t dx = (int32_t)xl - (int32_t)p.e0.x;                    // x - xlo
    const uint32_t den = (1u << ix.shift) + p.e1.x - p.e0.x;             // xhi - xlo in [1, 2^32)
    const uint32_t dy = p.e1.y - p.e0.y;                                 // yhi - ylo
    const double num = (double)(dx == 0 ? 1 : dx);                       // see interpolate(): 0 / d takes the slow path
    const double frac = __ddiv_rn(num, (double)den);
    const double rise = __dmul_rn((double)dy, frac);
    const double base = __dadd_rn(0.5, (double)p.e0.y);
    const double sum = __dadd_rn(base, rise);
    long long r = __double2ll_rz(sum);
    if (dx == 0) r = (long long)p.e0.y;
    if (r < 0) r = 0;
    return (uint64_t)r;
  }
  long long xhi, yhi;
  if (b + 1 == B) {
    xhi = ix.last_x;
    yhi = ix.last_y;
  } else {
    if (p.e1.x & kNarrowFill) return (uint64_t)p.e0.y;  // next bucket is a copy of this one: xlo == xhi (:105)
    xhi = (long long)(((b + 1) << ix.shift) + p.e1.x);
    yhi = (long long)p.e1.y;
  }
  long long xlo;
  if (p.e0.x & kNarrowFill) {
    const uint64_t src = b - (p.e0.x & ~kNarrowFill);
    xlo = (long long)((src << ix.shift) + ld_u32x2_pol(ix.narrow + src, pol).x);
  } else {
    xlo = (long long)((b << ix.shift) + p.e0.x);
  }
  return interpolate((long long)x, xlo, (long long)p.e0.y, xhi, yhi);
}

__device__ __forceinline__ uint64_t predict_rank(const IndexView& ix, uint64_t x, uint64_t pol) {
  const uint64_t b = x >> ix.shift;
  if (ix.narrow) return narrow_finish(ix, x, narrow_load(ix, x, pol), pol);
  const longlong2 lo = ld_s64x2_pol(reinterpret_cast<const longlong2*>(ix.model + b), pol);
  const longlong2 hi = ld_s64x2_pol(reinterpret_cast<const longlong2*>(ix.model + b + 1), pol);
  return interpolate((long long)x, lo.x, lo.y, hi.x, hi.y);
}

// Partitioned batches (partition.cu, query.cu kmer_query_ordered_kernel).
// For k <= 25 the 14-bit slot of a query rides in bits 50-63 of its partitioned k-mer word (2k + 14 <= 64): one shared-
// memory store, one global store and one load per query instead of two of each (the scatter is bound by exactly those:
// ncu r2k, stalls mio_throttle + lg_throttle).  The slot pointer then carries this tag instead of an array.
constexpr int kSlotShift = 50;
constexpr uint64_t kSlotKmerMask = (1ull << kSlotShift) - 1ull;
__host__ __device__ inline const uint16_t* slot_in_kmer_tag() { return reinterpret_cast<const uint16_t*>(uintptr_t(1)); }


__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace sb

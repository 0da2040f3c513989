// common.cuh -- shared definitions for the sapling_b200 CUDA library (sm_100a).
//
// Data layout in HBM:
//   genome : 2-bit packed, 32 bases per 64-bit word, base i in bits [63-2(i%32)-1, 63-2(i%32)]
//            of word i/32 ("big-endian in word": integer order of a word == lexicographic order
//            of its 32 bases).  A=0 C=1 G=2 T=3 (the vals[] table of the reference,
//            sapling_api.h:494-498).  Padded with GENOME_PAD_WORDS zero words.
//   sa     : uint32_t[n], rank -> text position (the reference's `rev`, sapling_api.h:41).
//   model  : ModelEntry[(1<<nb)+1] = {x, y} checkpoints (xlist/ylist, sapling_api.h:65) interleaved
//            so that the two checkpoints a query needs are 32 contiguous bytes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int GENOME_PAD_WORDS = 4;

struct __align__(16) ModelEntry {
  long long x;
  long long y;
};

// Everything the query kernels read; passed by value as a kernel argument (lives in the
// constant bank, so the scalars cost no memory traffic).
struct IndexView {
  const uint64_t* genome;
  const uint32_t* sa;
  const ModelEntry* model;
  uint64_t n;
  int k;
  int nb;
  int shift;  // 2k - nb
  int maxOver, maxUnder, mostOver, mostUnder;
  int compat;  // 1: the reference's (int)predicted window arithmetic (sapling_api.h:209,225)
  unsigned long long* oob_counter;  // incremented when predicted >= n (reference: UB)
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

__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p) { return __ldg(p); }

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

// queryPiecewiseLinear (sapling_api.h:98-109).  IEEE double, one rounding per operation, no FMA
// contraction, in the reference's evaluation order:
//   (long long)( (.5 + ylo) + (yhi - ylo) * ( ((x - xlo) * 1.) / (xhi - xlo) ) ), clamped at 0.
__device__ __forceinline__ uint64_t interpolate(long long x, long long xlo, long long ylo, long long xhi,
                                                long long yhi) {
  if (xlo == xhi) return (uint64_t)ylo;
  const double num = __ll2double_rn(x - xlo);
  const double frac = __ddiv_rn(num, __ll2double_rn(xhi - xlo));
  const double rise = __dmul_rn(__ll2double_rn(yhi - ylo), frac);
  const double base = __dadd_rn(0.5, __ll2double_rn(ylo));
  const double sum = __dadd_rn(base, rise);
  long long p = __double2ll_rz(sum);
  if (p < 0) p = 0;
  return (uint64_t)p;
}

__device__ __forceinline__ uint64_t predict_rank(const IndexView& ix, uint64_t x) {
  const uint64_t b = x >> ix.shift;
  const longlong2 lo = __ldg(reinterpret_cast<const longlong2*>(ix.model + b));
  const longlong2 hi = __ldg(reinterpret_cast<const longlong2*>(ix.model + b + 1));
  return interpolate((long long)x, lo.x, lo.y, hi.x, hi.y);
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace sb

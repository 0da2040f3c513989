// genome.cu -- packed-genome utilities and the LCP-derived side arrays.
//
// Everything the reference asks of its 8n-byte LCP array (sa.h:27) is "is lcp[r] >= k?"
// (krmq_init sa.h:33-43, countHits sapling_api.h:258,287), so the device keeps one byte flag per
// rank computed from SA + packed genome by a bounded compare; the full LCP is only produced when a
// reference-format .sa file has to be written.
#include "build.cuh"
#include "common.cuh"

namespace sb {

namespace {

inline int grid_for(uint64_t m, int per_sm = 16) {
  uint64_t g = (m + 255) / 256;
  if (g > 148ull * per_sm) g = 148ull * per_sm;
  if (g < 1) g = 1;
  return (int)g;
}

// A=0 C=1 G=2 T=3 from the ASCII code (valid for upper-case ACGT only)
__device__ __forceinline__ uint64_t ascii_code(unsigned c) { return ((c >> 1) ^ (c >> 2)) & 3u; }

__global__ void pack_kernel(const char* __restrict__ ascii, uint64_t n, uint64_t* __restrict__ packed,
                            uint64_t nwords) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    const uint64_t base = w * 32;
    uint64_t v = 0;
    if (base + 32 <= n) {
      const uint4* p = reinterpret_cast<const uint4*>(ascii + base);
      const uint4 a = __ldg(p), b = __ldg(p + 1);
      const unsigned ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int t = 0; t < 4; t++) v = (v << 2) | ascii_code((ws[j] >> (8 * t)) & 0xFFu);
      }
    } else {
      for (int j = 0; j < 32; j++) {
        const uint64_t i = base + j;
        v = (v << 2) | (i < n ? ascii_code((unsigned char)ascii[i]) : 0ull);
      }
    }
    packed[w] = v;
  }
}

__global__ void unpack_kernel(const uint64_t* __restrict__ packed, uint64_t n, char* __restrict__ ascii) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned c = (unsigned)(packed[i >> 5] >> (62 - 2 * (i & 31))) & 3u;
    ascii[i] = "ACGT"[c];
  }
}

__global__ void synth_kernel(uint64_t seed, uint64_t n, uint64_t* __restrict__ packed, uint64_t nwords) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    const uint64_t base = w * 32;
    uint64_t v = 0;
    for (int j = 0; j < 32; j++) {
      const uint64_t i = base + j;
      v = (v << 2) | (i < n ? (splitmix64(seed + i) >> 62) : 0ull);
    }
    packed[w] = v;
  }
}

// LCP of the suffixes at text positions a and b, capped at `cap`
__device__ __forceinline__ uint64_t suffix_lcp(const uint64_t* __restrict__ genome, uint64_t n, uint64_t a,
                                               uint64_t b, uint64_t cap) {
  uint64_t lim = n - a < n - b ? n - a : n - b;
  if (cap < lim) lim = cap;
  uint64_t l = 0;
  while (l < lim) {
    const uint64_t x = load_bases32(genome, a + l) ^ load_bases32(genome, b + l);
    if (x) {
      l += (uint64_t)(__clzll((long long)x) >> 1);
      break;
    }
    l += 32;
  }
  return l < lim ? l : lim;
}

__device__ __forceinline__ unsigned base_at(const uint64_t* __restrict__ genome, uint64_t i) {
  return (unsigned)(__ldg(genome + (i >> 5)) >> (62 - 2 * (i & 31))) & 3u;
}

__global__ void lcp_kernel(const uint64_t* __restrict__ genome, uint64_t n, const uint32_t* __restrict__ sa,
                           uint32_t* __restrict__ lcp) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r + 1 < n; r += stride)
    lcp[r] = (uint32_t)suffix_lcp(genome, n, sa[r], sa[r + 1], ~0ull);
}

__global__ void kflag_kernel(const uint64_t* __restrict__ genome, uint64_t n, const uint32_t* __restrict__ sa,
                             uint32_t k, uint8_t* __restrict__ kflag) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
    kflag[r] = (r + 1 < n && suffix_lcp(genome, n, sa[r], sa[r + 1], k) >= k) ? 1 : 0;
}

__global__ void check_kernel(const uint64_t* __restrict__ genome, uint64_t n, const uint32_t* __restrict__ sa,
                             const uint32_t* __restrict__ isa, uint32_t max_chars,
                             unsigned long long* __restrict__ counters) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0, und = 0, perm = 0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const uint64_t a = sa[r];
    if (a >= n || (isa && isa[a] != (uint32_t)r)) perm++;
    if (r + 1 < n && a < n) {
      const uint64_t b = sa[r + 1];
      if (b >= n) continue;
      const uint64_t l = suffix_lcp(genome, n, a, b, max_chars);
      if (a + l == n) continue;  // a is a proper prefix of b: in order
      if (b + l == n) { bad++; continue; }
      if (l == max_chars) { und++; continue; }
      if (base_at(genome, a + l) >= base_at(genome, b + l)) bad++;
    }
  }
  if (bad) atomicAdd(counters + 0, bad);
  if (und) atomicAdd(counters + 1, und);
  if (perm) atomicAdd(counters + 2, perm);
}

__global__ void count_hits_kernel(const uint8_t* __restrict__ kflag, uint64_t n, uint32_t k,
                                  const uint32_t* __restrict__ sa_pos, size_t count, uint32_t maxHits,
                                  uint32_t* __restrict__ left, uint32_t* __restrict__ right) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    const uint64_t p = sa_pos[t];
    uint32_t i = 0;
    // countHitsRight, sapling_api.h:254-263
    for (; i < maxHits; i++)
      if ((uint64_t)i + p > n - k || !kflag[i + p]) break;
    right[t] = i;
    // countHitsLeft, sapling_api.h:283-289 (note: starts at lcp[sa_pos] itself, as the reference does)
    for (i = 0; i < maxHits; i++)
      if (p < i || !kflag[p - i]) break;
    left[t] = i;
  }
}

}  // namespace

int pack_genome(const char* d_ascii, uint64_t n, uint64_t* d_packed, cudaStream_t st) {
  const uint64_t nw = packed_words(n);
  pack_kernel<<<grid_for(nw), 256, 0, st>>>(d_ascii, n, d_packed, nw);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int unpack_genome(const uint64_t* d_packed, uint64_t n, char* d_ascii, cudaStream_t st) {
  unpack_kernel<<<grid_for(n), 256, 0, st>>>(d_packed, n, d_ascii);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int synth_genome_packed(uint64_t seed, uint64_t n, uint64_t* d_packed, cudaStream_t st) {
  const uint64_t nw = packed_words(n);
  synth_kernel<<<grid_for(nw), 256, 0, st>>>(seed, n, d_packed, nw);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int compute_lcp(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, uint32_t* d_lcp, cudaStream_t st) {
  lcp_kernel<<<grid_for(n), 256, 0, st>>>(d_genome, n, d_sa, d_lcp);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int compute_kflags(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, int k, uint8_t* d_kflag,
                   cudaStream_t st) {
  kflag_kernel<<<grid_for(n), 256, 0, st>>>(d_genome, n, d_sa, (uint32_t)k, d_kflag);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int check_suffix_array(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                       uint32_t max_chars, uint64_t* bad_order, uint64_t* undecided, uint64_t* bad_perm,
                       cudaStream_t st) {
  unsigned long long* d_c = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_c, 3 * sizeof(unsigned long long)));
  SB_CUDA_CHECK(cudaMemsetAsync(d_c, 0, 3 * sizeof(unsigned long long), st));
  check_kernel<<<grid_for(n), 256, 0, st>>>(d_genome, n, d_sa, d_isa, max_chars, d_c);
  unsigned long long h[3] = {0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d_c, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_c);
  SB_CUDA_CHECK(e);
  if (bad_order) *bad_order = h[0];
  if (undecided) *undecided = h[1];
  if (bad_perm) *bad_perm = h[2];
  return 0;
}

int count_hits(const uint8_t* d_kflag, uint64_t n, int k, const uint32_t* d_sa_pos, size_t count,
               uint32_t maxHits, uint32_t* d_left, uint32_t* d_right, cudaStream_t st) {
  count_hits_kernel<<<grid_for(count), 256, 0, st>>>(d_kflag, n, (uint32_t)k, d_sa_pos, count, maxHits, d_left,
                                                     d_right);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sb

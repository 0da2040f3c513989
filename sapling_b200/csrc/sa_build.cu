// sa_build.cu -- suffix array construction on the GPU (replaces the reference's host builders:
// DC3 in src/sa.h:82-183 and the libdivsufsort pipeline suffixarray/refToSuffixArray.sh).
//
// The suffix array of a text is unique, so any correct builder reproduces the reference's `rev`
// (and, inverted, the ISA stored in the .sa file).  Method: one 64-bit radix sort of every suffix
// by its first 27 characters over the alphabet {$<A<C<G<T} (so a suffix that runs into the end of
// the text sorts before its extensions, as in sa.h:18), then prefix doubling restricted to the
// suffixes that are still tied.  For a random genome the first sort resolves all but a handful of
// suffixes; repetitive genomes take ~log2(longest repeat / 27) extra rounds over the tied subset.
//
// The radix sorts are cub::DeviceRadixSort (library code, off the query hot path).
#include <cub/cub.cuh>

#include "build.cuh"
#include "common.cuh"

namespace sb {

namespace {

constexpr int kKeyChars = 27;  // 5^27 < 2^63

__global__ void sa_init_keys(const uint64_t* __restrict__ genome, uint64_t n, uint64_t* __restrict__ keys,
                             uint32_t* __restrict__ pos) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t w = load_bases32(genome, i);
    const uint64_t room = n - i;
    uint64_t key = 0;
#pragma unroll
    for (int j = 0; j < kKeyChars; j++) {
      const uint64_t c = ((uint64_t)j < room) ? (((w >> (62 - 2 * j)) & 3ull) + 1ull) : 0ull;
      key = key * 5ull + c;
    }
    keys[i] = key;
    pos[i] = (uint32_t)i;
  }
}

// head[i] = 1 when slot i starts a group of equal keys
__global__ void sa_mark_heads(const uint64_t* __restrict__ keys, uint64_t m, uint8_t* __restrict__ head) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// value fed to the max-scan: the slot index of a head, 0 otherwise
__global__ void sa_head_slots(const uint8_t* __restrict__ head, const uint32_t* __restrict__ slot, uint64_t m,
                              uint32_t* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    out[i] = head[i] ? (slot ? slot[i] : (uint32_t)i) : 0u;
}

// rank[pos[i]] = slot of the head of i's group; tied[i] = 1 when the group has more than one member
__global__ void sa_assign_ranks(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ headslot,
                                const uint8_t* __restrict__ head, uint64_t m, uint32_t* __restrict__ rank,
                                uint8_t* __restrict__ tied) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    rank[pos[i]] = headslot[i];
    const bool single = head[i] && (i + 1 == m || head[i + 1]);
    tied[i] = single ? 0 : 1;
  }
}

__global__ void sa_iota(uint32_t* __restrict__ a, uint64_t m) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) a[i] = (uint32_t)i;
}

// doubling key of a tied suffix: (rank of its group) . (rank of the suffix h characters further + 1,
// or 0 when that runs past the end; such a suffix is never tied with another, see header)
__global__ void sa_doubling_keys(const uint32_t* __restrict__ slot, const uint32_t* __restrict__ sa,
                                 const uint32_t* __restrict__ rank, uint64_t m, uint64_t n, uint64_t h,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ pos) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += stride) {
    const uint32_t p = sa[slot[t]];
    const uint64_t q = (uint64_t)p + h;
    const uint64_t second = q < n ? (uint64_t)rank[q] + 1ull : 0ull;
    keys[t] = ((uint64_t)rank[p] << 32) | second;  // rank <= n-1 and second <= n both fit 32 bits
    pos[t] = p;
  }
}

__global__ void sa_scatter_back(const uint32_t* __restrict__ slot, const uint32_t* __restrict__ pos, uint64_t m,
                                uint32_t* __restrict__ sa) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += stride) sa[slot[t]] = pos[t];
}

__global__ void sa_invert(const uint32_t* __restrict__ src, uint64_t n, uint32_t* __restrict__ dst) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[src[i]] = (uint32_t)i;
}

// rank lines (common.cuh IndexView): one thread per 32-byte sector
__global__ void rank_lines_kernel(const uint64_t* __restrict__ genome, const uint32_t* __restrict__ sa, uint64_t n,
                                  int bases, uint64_t sectors, uint32_t* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < sectors; s += stride) {
    const uint64_t r0 = s * 4;
    uint32_t v[8];
    pack_rank_sector(genome, sa, n, bases, r0, v);
    uint4* dst = reinterpret_cast<uint4*>(out + s * 8);
    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

inline int grid_for(uint64_t m) {
  uint64_t g = (m + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  if (g < 1) g = 1;
  return (int)g;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

int invert_permutation(const uint32_t* d_src, uint64_t n, uint32_t* d_dst, cudaStream_t st) {
  sa_invert<<<grid_for(n), 256, 0, st>>>(d_src, n, d_dst);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int build_rank_lines(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, int bases, uint32_t* d_lines,
                     cudaStream_t st) {
  const uint64_t sectors = line_sectors(n);
  rank_lines_kernel<<<grid_for(sectors), 256, 0, st>>>(d_genome, d_sa, n, bases, sectors, d_lines);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// d_sa (out): rank -> position.  d_isa (out): position -> rank.  Both uint32[n], caller-allocated.
int build_suffix_array(const uint64_t* d_genome, uint64_t n, uint32_t* d_sa, uint32_t* d_isa, cudaStream_t st,
                       int* rounds_out) {
  if (n == 0 || n >= 0xFFFFFF00ull) {
    set_error("build_suffix_array: n=%llu unsupported (need 0 < n < 2^32-256)", (unsigned long long)n);
    return -1;
  }
  DevBuf keysA, keysB, posB, head, tied, headslot;
  SB_CUDA_CHECK(keysA.alloc(n * 8));
  SB_CUDA_CHECK(keysB.alloc(n * 8));
  SB_CUDA_CHECK(posB.alloc(n * 4));
  SB_CUDA_CHECK(head.alloc(n));
  SB_CUDA_CHECK(tied.alloc(n));
  SB_CUDA_CHECK(headslot.alloc(n * 4));

  sa_init_keys<<<grid_for(n), 256, 0, st>>>(d_genome, n, keysA.as<uint64_t>(), d_sa);
  SB_CUDA_CHECK(cudaGetLastError());

  cub::DoubleBuffer<uint64_t> dk(keysA.as<uint64_t>(), keysB.as<uint64_t>());
  cub::DoubleBuffer<uint32_t> dv(d_sa, posB.as<uint32_t>());
  size_t tmp_bytes = 0;
  SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (unsigned long long)n, 0, 63, st));
  size_t scan_bytes = 0;
  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, headslot.as<uint32_t>(), headslot.as<uint32_t>(),
                                               MaxOp(), (unsigned long long)n, st));
  size_t sel_bytes = 0;
  SB_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, sel_bytes, (uint32_t*)nullptr, (uint8_t*)nullptr,
                                           (uint32_t*)nullptr, (unsigned long long*)nullptr,
                                           (unsigned long long)n, st));
  DevBuf tmp;
  size_t tb = tmp_bytes > scan_bytes ? tmp_bytes : scan_bytes;
  if (sel_bytes > tb) tb = sel_bytes;
  SB_CUDA_CHECK(tmp.alloc(tb));
  SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (unsigned long long)n, 0, 63, st));
  if (dv.Current() != d_sa)
    SB_CUDA_CHECK(cudaMemcpyAsync(d_sa, dv.Current(), n * 4, cudaMemcpyDeviceToDevice, st));
  uint64_t* keys_sorted = dk.Current();
  uint64_t* keys_other = dk.Alternate();

  // group structure after the first sort
  sa_mark_heads<<<grid_for(n), 256, 0, st>>>(keys_sorted, n, head.as<uint8_t>());
  sa_head_slots<<<grid_for(n), 256, 0, st>>>(head.as<uint8_t>(), nullptr, n, headslot.as<uint32_t>());
  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, scan_bytes, headslot.as<uint32_t>(), headslot.as<uint32_t>(),
                                               MaxOp(), (unsigned long long)n, st));
  sa_assign_ranks<<<grid_for(n), 256, 0, st>>>(d_sa, headslot.as<uint32_t>(), head.as<uint8_t>(), n, d_isa,
                                               tied.as<uint8_t>());
  SB_CUDA_CHECK(cudaGetLastError());

  // tied subset: slots (in increasing order) whose group has > 1 member
  DevBuf count;
  SB_CUDA_CHECK(count.alloc(8));
  // reuse keys_other (8n bytes) as the iota + slot storage to keep the footprint down
  uint32_t* iota = reinterpret_cast<uint32_t*>(keys_other);
  uint32_t* slots = iota + n;  // second half of the 8n-byte buffer
  sa_iota<<<grid_for(n), 256, 0, st>>>(iota, n);
  SB_CUDA_CHECK(cub::DeviceSelect::Flagged(tmp.p, sel_bytes, iota, tied.as<uint8_t>(), slots,
                                           count.as<unsigned long long>(), (unsigned long long)n, st));
  unsigned long long m = 0;
  SB_CUDA_CHECK(cudaMemcpyAsync(&m, count.p, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA_CHECK(cudaStreamSynchronize(st));

  int rounds = 0;
  if (m > 0) {
    // subset buffers (m only shrinks from here on)
    DevBuf sk1, sk2, sp1, sp2, sl1, sl2, shead, stied, shs;
    SB_CUDA_CHECK(sk1.alloc(m * 8));
    SB_CUDA_CHECK(sk2.alloc(m * 8));
    SB_CUDA_CHECK(sp1.alloc(m * 4));
    SB_CUDA_CHECK(sp2.alloc(m * 4));
    SB_CUDA_CHECK(sl1.alloc(m * 4));
    SB_CUDA_CHECK(sl2.alloc(m * 4));
    SB_CUDA_CHECK(shead.alloc(m));
    SB_CUDA_CHECK(stied.alloc(m));
    SB_CUDA_CHECK(shs.alloc(m * 4));
    SB_CUDA_CHECK(cudaMemcpyAsync(sl1.p, slots, m * 4, cudaMemcpyDeviceToDevice, st));
    uint32_t* cur_slots = sl1.as<uint32_t>();
    uint32_t* alt_slots = sl2.as<uint32_t>();
    uint64_t h = kKeyChars;
    while (m > 0) {
      rounds++;
      sa_doubling_keys<<<grid_for(m), 256, 0, st>>>(cur_slots, d_sa, d_isa, m, n, h, sk1.as<uint64_t>(),
                                                    sp1.as<uint32_t>());
      cub::DoubleBuffer<uint64_t> k2(sk1.as<uint64_t>(), sk2.as<uint64_t>());
      cub::DoubleBuffer<uint32_t> v2(sp1.as<uint32_t>(), sp2.as<uint32_t>());
      size_t need = 0;
      SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, need, k2, v2, (unsigned long long)m, 0, 64, st));
      if (need > tb) { set_error("build_suffix_array: temp storage too small"); return -1; }
      SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, need, k2, v2, (unsigned long long)m, 0, 64, st));
      sa_scatter_back<<<grid_for(m), 256, 0, st>>>(cur_slots, v2.Current(), m, d_sa);
      sa_mark_heads<<<grid_for(m), 256, 0, st>>>(k2.Current(), m, shead.as<uint8_t>());
      sa_head_slots<<<grid_for(m), 256, 0, st>>>(shead.as<uint8_t>(), cur_slots, m, shs.as<uint32_t>());
      size_t need2 = 0;
      SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, need2, shs.as<uint32_t>(), shs.as<uint32_t>(), MaxOp(),
                                                   (unsigned long long)m, st));
      if (need2 > tb) { set_error("build_suffix_array: temp storage too small"); return -1; }
      SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, need2, shs.as<uint32_t>(), shs.as<uint32_t>(), MaxOp(),
                                                   (unsigned long long)m, st));
      // ranks change only after every key of this round has been formed (separate kernel)
      sa_assign_ranks<<<grid_for(m), 256, 0, st>>>(v2.Current(), shs.as<uint32_t>(), shead.as<uint8_t>(), m, d_isa,
                                                   stied.as<uint8_t>());
      size_t need3 = 0;
      SB_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, need3, cur_slots, stied.as<uint8_t>(), alt_slots,
                                               count.as<unsigned long long>(), (unsigned long long)m, st));
      if (need3 > tb) { set_error("build_suffix_array: temp storage too small"); return -1; }
      SB_CUDA_CHECK(cub::DeviceSelect::Flagged(tmp.p, need3, cur_slots, stied.as<uint8_t>(), alt_slots,
                                               count.as<unsigned long long>(), (unsigned long long)m, st));
      SB_CUDA_CHECK(cudaMemcpyAsync(&m, count.p, 8, cudaMemcpyDeviceToHost, st));
      SB_CUDA_CHECK(cudaStreamSynchronize(st));
      uint32_t* t = cur_slots; cur_slots = alt_slots; alt_slots = t;
      h *= 2;
      if (rounds > 40) { set_error("build_suffix_array: did not converge"); return -1; }
    }
  }
  if (rounds_out) *rounds_out = rounds;
  SB_CUDA_CHECK(cudaGetLastError());
  SB_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace sb

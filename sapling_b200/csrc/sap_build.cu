// sap_build.cu -- the .sap build step on the GPU: Sapling::buildPiecewiseLinear, getError and
// errorStats (reference sapling_api.h:384-487, 309-337, 342-379; sa.h:33-57).
//
// The reference makes two sequential passes over the text with a rolling hash.  Restated as
// data-parallel steps with the same outcome:
//   1. checkpoint b = smallest k-mer value seen in bucket b, y = SA rank of its FIRST text
//      occurrence (strict '>' at :422)                       -> two atomicMin passes per bucket
//      last checkpoint = largest k-mer overall, first occurrence (:429-433)
//   2. empty buckets copy their left neighbour (:437-449)     -> max-scan of "last non-empty"
//   3. signed error of every k-mer (:458-481); an under-prediction slides y right along the run
//      of ranks sharing the k-prefix (:311-323), i.e. y' = min(predict, y + krmqb[y]); the
//      over-prediction shift is computed and dropped by the reference (:325-336)
//                                                            -> run ends by a reverse min-scan
//   4. max / sum / counts by reduction; the 95th-percentile bounds (:370-376) by a full radix
//      sort of the signed errors (exact order statistic).
// The scans and sorts are cub:: primitives (library code, off the query hot path).
#include <cub/cub.cuh>

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

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

__device__ __forceinline__ uint64_t kmer_at(const uint64_t* __restrict__ genome, uint64_t i, int k) {
  return load_bases_upto(genome, i, (unsigned)k) >> (64 - 2 * k);
}

constexpr unsigned long long kEmpty = ~0ull;

__global__ void fill_u64(unsigned long long* a, uint64_t m, unsigned long long v) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) a[i] = v;
}
__global__ void fill_u32(uint32_t* a, uint64_t m, uint32_t v) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) a[i] = v;
}

// step 1a: per-bucket minimum k-mer value; global maximum in xmax[0]
__global__ void bucket_min_x(const uint64_t* __restrict__ genome, uint64_t nk, int k, int shift,
                             unsigned long long* __restrict__ xmin, unsigned long long* __restrict__ xmax) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mx = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += stride) {
    const unsigned long long x = kmer_at(genome, i, k);
    atomicMin(xmin + (x >> shift), x);
    mx = x > mx ? x : mx;
  }
  // warp max then one atomic
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = t > mx ? t : mx;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(xmax, mx);
}

// step 1b: first text occurrence of each bucket minimum / of the global maximum
__global__ void bucket_min_pos(const uint64_t* __restrict__ genome, uint64_t nk, int k, int shift,
                               const unsigned long long* __restrict__ xmin,
                               const unsigned long long* __restrict__ xmax, uint32_t* __restrict__ posmin,
                               uint32_t* __restrict__ posmax) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const unsigned long long gmax = *xmax;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += stride) {
    const unsigned long long x = kmer_at(genome, i, k);
    const uint64_t b = x >> shift;
    if (x == xmin[b]) atomicMin(posmin + b, (uint32_t)i);
    if (x == gmax) atomicMin(posmax, (uint32_t)i);
  }
}

// scan input: b+1 for a non-empty bucket, 0 for an empty one
__global__ void nonempty_marks(const unsigned long long* __restrict__ xmin, uint64_t B, uint32_t* __restrict__ mark) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride)
    mark[b] = xmin[b] != kEmpty ? (uint32_t)(b + 1) : 0u;
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
struct MinOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; }
};

// steps 1c+2: write the checkpoints, forward-filling empty buckets
__global__ void write_model(const unsigned long long* __restrict__ xmin, const uint32_t* __restrict__ posmin,
                            const uint32_t* __restrict__ src, const uint32_t* __restrict__ isa, uint64_t B,
                            const unsigned long long* __restrict__ xmax, const uint32_t* __restrict__ posmax,
                            ModelEntry* __restrict__ model) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += stride) {
    ModelEntry e;
    if (b == B) {
      e.x = (long long)*xmax;
      e.y = (long long)isa[*posmax];
    } else if (src[b] == 0) {
      e.x = 0;  // :437-441 and the fill loop copying it rightwards
      e.y = 0;
    } else {
      const uint64_t s = src[b] - 1;
      e.x = (long long)xmin[s];
      e.y = (long long)isa[posmin[s]];
    }
    model[b] = e;
  }
}

// reverse-scan input: rev[n-1-r] = r when the k-prefix run breaks after rank r, else "infinity"
__global__ void runbreak_marks(const uint8_t* __restrict__ kflag, uint64_t n, uint32_t* __restrict__ rev) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
    rev[n - 1 - r] = kflag[r] ? 0xFFFFFFFFu : (uint32_t)r;
}

// step 3: val = getError(inv[i], queryPiecewiseLinear(x_i))
__global__ void error_pass(const uint64_t* __restrict__ genome, uint64_t n, uint64_t nk, int k, int shift,
                           const ModelEntry* __restrict__ model, const uint32_t* __restrict__ isa,
                           const uint32_t* __restrict__ runend_rev, int* __restrict__ val,
                           long long* __restrict__ dump) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += stride) {
    const uint64_t x = kmer_at(genome, i, k);
    const uint64_t b = x >> shift;
    const ModelEntry lo = model[b], hi = model[b + 1];
    const uint64_t predict = interpolate((long long)x, lo.x, lo.y, hi.x, hi.y);
    uint64_t y = isa[i];
    const uint64_t y0 = y;
    if (y < predict) {
      // :311-323: largest y' in [y, predict] with ranks y..y' sharing the k-prefix
      const uint64_t end = runend_rev[n - 1 - y];  // y + krmqb[y]
      y = end < predict ? end : predict;
    }
    const int v = (int)((long long)y - (long long)predict);
    val[i] = v;
    if (dump) {
      dump[3 * i + 0] = (long long)y0;
      dump[3 * i + 1] = (long long)predict;
      dump[3 * i + 2] = (long long)v;
    }
  }
}

// step 4a: counters[0]=maxOver [1]=maxUnder [2]=sum|val| [3]=nOver [4]=nUnder [5]=perfect
__global__ void error_reduce(const int* __restrict__ val, uint64_t nk, unsigned long long* __restrict__ counters) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mo = 0, mu = 0, tot = 0, no = 0, nu = 0, pf = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += stride) {
    const long long v = val[i];
    if (v > 0) { no++; tot += (unsigned long long)v; if ((unsigned long long)v > mo) mo = (unsigned long long)v; }
    else if (v < 0) { nu++; tot += (unsigned long long)(-v); if ((unsigned long long)(-v) > mu) mu = (unsigned long long)(-v); }
    else pf++;
  }
  for (int o = 16; o; o >>= 1) {
    unsigned long long t;
    t = __shfl_xor_sync(0xffffffffu, mo, o); mo = t > mo ? t : mo;
    t = __shfl_xor_sync(0xffffffffu, mu, o); mu = t > mu ? t : mu;
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    no += __shfl_xor_sync(0xffffffffu, no, o);
    nu += __shfl_xor_sync(0xffffffffu, nu, o);
    pf += __shfl_xor_sync(0xffffffffu, pf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(counters + 0, mo);
    atomicMax(counters + 1, mu);
    atomicAdd(counters + 2, tot);
    atomicAdd(counters + 3, no);
    atomicAdd(counters + 4, nu);
    atomicAdd(counters + 5, pf);
  }
}

}  // namespace

int build_model(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                const uint8_t* d_kflag, int k, int nb, ModelEntry* d_model, ModelStats* stats,
                int64_t* h_errdump, cudaStream_t st) {
  (void)d_sa;
  if (n < (uint64_t)k) { set_error("build_model: genome shorter than k"); return -1; }
  if (nb < 1 || nb > 2 * k || nb > 31) { set_error("build_model: need 1 <= nb <= min(2k,31), got %d", nb); return -1; }
  const uint64_t nk = n - (uint64_t)k + 1;
  const uint64_t B = 1ull << nb;
  const int shift = 2 * k - nb;

  DevBuf xmin, posmin, mark, scal, tmp;
  SB_CUDA_CHECK(xmin.alloc((B + 1) * 8));
  SB_CUDA_CHECK(posmin.alloc((B + 1) * 4));
  SB_CUDA_CHECK(mark.alloc(B * 4));
  SB_CUDA_CHECK(scal.alloc(64));
  unsigned long long* d_xmax = scal.as<unsigned long long>();
  uint32_t* d_posmax = reinterpret_cast<uint32_t*>(d_xmax + 1);
  fill_u64<<<grid_for(B + 1), 256, 0, st>>>(xmin.as<unsigned long long>(), B + 1, kEmpty);
  fill_u32<<<grid_for(B + 1), 256, 0, st>>>(posmin.as<uint32_t>(), B + 1, 0xFFFFFFFFu);
  SB_CUDA_CHECK(cudaMemsetAsync(d_xmax, 0, 8, st));
  SB_CUDA_CHECK(cudaMemsetAsync(d_posmax, 0xFF, 4, st));
  bucket_min_x<<<grid_for(nk), 256, 0, st>>>(d_genome, nk, k, shift, xmin.as<unsigned long long>(), d_xmax);
  bucket_min_pos<<<grid_for(nk), 256, 0, st>>>(d_genome, nk, k, shift, xmin.as<unsigned long long>(), d_xmax,
                                               posmin.as<uint32_t>(), d_posmax);
  nonempty_marks<<<grid_for(B), 256, 0, st>>>(xmin.as<unsigned long long>(), B, mark.as<uint32_t>());
  SB_CUDA_CHECK(cudaGetLastError());

  // temp storage sized for the largest primitive used below
  DevBuf runrev, val, val2;
  SB_CUDA_CHECK(runrev.alloc(n * 4));
  SB_CUDA_CHECK(val.alloc(nk * 4));
  SB_CUDA_CHECK(val2.alloc(nk * 4));
  size_t s1 = 0, s2 = 0, s3 = 0;
  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, s1, mark.as<uint32_t>(), mark.as<uint32_t>(), MaxOp(),
                                               (unsigned long long)B, st));
  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(nullptr, s2, runrev.as<uint32_t>(), runrev.as<uint32_t>(), MinOp(),
                                               (unsigned long long)n, st));
  cub::DoubleBuffer<int> dv(val.as<int>(), val2.as<int>());
  SB_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, s3, dv, (unsigned long long)nk, 0, 32, st));
  size_t tb = s1 > s2 ? s1 : s2;
  if (s3 > tb) tb = s3;
  SB_CUDA_CHECK(tmp.alloc(tb));

  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, s1, mark.as<uint32_t>(), mark.as<uint32_t>(), MaxOp(),
                                               (unsigned long long)B, st));
  write_model<<<grid_for(B + 1), 256, 0, st>>>(xmin.as<unsigned long long>(), posmin.as<uint32_t>(),
                                               mark.as<uint32_t>(), d_isa, B, d_xmax, d_posmax, d_model);
  SB_CUDA_CHECK(cudaGetLastError());

  runbreak_marks<<<grid_for(n), 256, 0, st>>>(d_kflag, n, runrev.as<uint32_t>());
  SB_CUDA_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, s2, runrev.as<uint32_t>(), runrev.as<uint32_t>(), MinOp(),
                                               (unsigned long long)n, st));
  DevBuf dump;
  if (h_errdump) SB_CUDA_CHECK(dump.alloc(nk * 3 * 8));
  error_pass<<<grid_for(nk), 256, 0, st>>>(d_genome, n, nk, k, shift, d_model, d_isa, runrev.as<uint32_t>(),
                                           val.as<int>(), h_errdump ? dump.as<long long>() : nullptr);
  SB_CUDA_CHECK(cudaGetLastError());
  if (h_errdump)
    SB_CUDA_CHECK(cudaMemcpyAsync(h_errdump, dump.p, nk * 3 * 8, cudaMemcpyDeviceToHost, st));

  DevBuf counters;
  SB_CUDA_CHECK(counters.alloc(6 * 8));
  SB_CUDA_CHECK(cudaMemsetAsync(counters.p, 0, 6 * 8, st));
  error_reduce<<<grid_for(nk), 256, 0, st>>>(val.as<int>(), nk, counters.as<unsigned long long>());
  unsigned long long c[6];
  SB_CUDA_CHECK(cudaMemcpyAsync(c, counters.p, sizeof(c), cudaMemcpyDeviceToHost, st));
  SB_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, s3, dv, (unsigned long long)nk, 0, 32, st));
  SB_CUDA_CHECK(cudaStreamSynchronize(st));

  // errorStats (:342-379)
  const uint64_t nOver = c[3], nUnder = c[4], perfect = c[5];
  int maxOver = (int)c[0], maxUnder = (int)c[1];
  if (maxUnder < 2) maxUnder = 2;
  if (maxOver < 2) maxOver = 2;
  const uint64_t cnt = nOver + nUnder + perfect;
  const int meanError = (int)(.5 + (double)(c[2] / cnt));
  const double mostThreshold = 0.95;  // :35
  int mostOver = 0, mostUnder = 0;
  const int* sorted = dv.Current();  // ascending: unders (most negative first), zeros, overs
  if (nOver > 0) {
    const uint64_t j = (uint64_t)(mostThreshold * (double)nOver);
    SB_CUDA_CHECK(cudaMemcpy(&mostOver, sorted + (nUnder + perfect + j), 4, cudaMemcpyDeviceToHost));
  }
  if (nUnder > 0) {
    const uint64_t j = (uint64_t)(mostThreshold * (double)nUnder);
    int v = 0;
    SB_CUDA_CHECK(cudaMemcpy(&v, sorted + (nUnder - 1 - j), 4, cudaMemcpyDeviceToHost));
    mostUnder = -v;
  }
  if (mostOver < 1) mostOver = 1;
  if (mostUnder < 1) mostUnder = 1;
  if (stats) {
    stats->maxOver = maxOver; stats->maxUnder = maxUnder; stats->meanError = meanError;
    stats->mostOver = mostOver; stats->mostUnder = mostUnder;
    stats->perfect = perfect; stats->nOver = nOver; stats->nUnder = nUnder;
  }
  return 0;
}

}  // namespace sb

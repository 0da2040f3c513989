// model.cu -- the narrow (8 bytes per bucket) device layout of the piecewise-linear model.
//
// The reference keeps two int64 arrays xlist/ylist of (1<<nb)+1 checkpoints (sapling_api.h:65,
// 406-407): 16 bytes per bucket, 134 MB at the 100 Mbp default (nb=23), 4.3 GB at 3.1 Gbp (nb=28).
// Every checkpoint x lies inside its own bucket [b<<shift, (b+1)<<shift) unless the bucket is empty,
// in which case it is a copy of the nearest non-empty bucket to the left (:437-449); y < n < 2^32.
// So {x - (b<<shift) : 31 bits + 1 flag bit, y : 32 bits} reconstructs both values exactly, halves the
// table (67 MB at nb=23: it fits in the B200's 126 MB L2 next to the 25 MB packed genome) and puts the
// two checkpoints a query needs in one 32-byte sector three times out of four.
#include "build.cuh"
#include "common.cuh"

namespace sb {

namespace {

__global__ void narrow_kernel(const ModelEntry* __restrict__ model, uint64_t B, int shift, uint2* __restrict__ out,
                              int* __restrict__ bad) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += stride) {
    if (b == B) {  // pad entry: lets the query kernels read bucket b + 1 unconditionally; flagged, so never interpolated with
      out[b] = make_uint2(kNarrowFill, 0u);
      continue;
    }
    const long long x = model[b].x, y = model[b].y;
    const long long base = (long long)(b << shift);
    uint2 e;
    bool ok = x >= 0 && y >= 0 && y <= 0xFFFFFFFFll;
    if (ok && x >= base && x < base + (1ll << shift)) {
      e.x = (uint32_t)(x - base);
    } else if (ok && x < base) {
      // forward-filled copy of the bucket x belongs to
      const uint64_t src = (uint64_t)x >> shift;
      ok = model[src].x == x && model[src].y == y && (b - src) < 0x7FFFFFFFull;
      e.x = kNarrowFill | (uint32_t)(b - src);
    } else {
      ok = false;
      e.x = 0;
    }
    e.y = (uint32_t)y;
    out[b] = e;
    if (!ok) atomicExch(bad, 1);
  }
}

// the inverse: checkpoints first .. first+count-1 of xlist / ylist (sapling_api.h:65) rebuilt from the narrow table
__global__ void widen_kernel(const uint2* __restrict__ narrow, uint64_t B, int shift, long long last_x, long long last_y,
                             uint64_t first, uint64_t count, long long* __restrict__ xs, long long* __restrict__ ys) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const uint64_t b = first + i;
    long long x = last_x, y = last_y;
    if (b < B) {
      const uint2 e = narrow[b];
      const uint64_t src = (e.x & kNarrowFill) ? b - (e.x & ~kNarrowFill) : b;
      const uint32_t xoff = (e.x & kNarrowFill) ? narrow[src].x : e.x;
      x = (long long)((src << shift) + xoff);
      y = (long long)e.y;
    }
    xs[i] = x;
    ys[i] = y;
  }
}

}  // namespace

int widen_model(const uint2* d_narrow, int nb, int shift, long long last_x, long long last_y, uint64_t first,
                uint64_t count, long long* d_xs, long long* d_ys, cudaStream_t st) {
  if (count == 0) return 0;
  uint64_t g = (count + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  widen_kernel<<<(int)g, 256, 0, st>>>(d_narrow, 1ull << nb, shift, last_x, last_y, first, count, d_xs, d_ys);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// d_narrow holds (1 << nb) + 1 entries (the last one is a pad).
// Returns 0 and sets *ok = 1 when every checkpoint is representable (always true for a model built by
// buildPiecewiseLinear; a hand-edited .sap file may not be, then the wide table is used).
int build_narrow_model(const ModelEntry* d_model, int nb, int shift, uint2* d_narrow, int* ok, cudaStream_t st) {
  *ok = 0;
  if (shift < 0 || shift > 31) return 0;
  int* d_bad = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_bad, sizeof(int)));
  SB_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  const uint64_t B = 1ull << nb;
  uint64_t g = (B + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  narrow_kernel<<<(int)g, 256, 0, st>>>(d_model, B, shift, d_narrow, d_bad);
  int bad = 1;
  cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_bad);
  SB_CUDA_CHECK(e);
  *ok = bad ? 0 : 1;
  return 0;
}

}  // namespace sb

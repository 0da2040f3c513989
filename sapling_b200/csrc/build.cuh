// build.cuh -- index construction entry points (host functions launching CUDA work).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb {

// ---- genome ----
// ASCII A/C/G/T (device) -> 2-bit packed words.  d_packed must hold packed_words(n) words.
inline uint64_t packed_words(uint64_t n) { return (n + 31) / 32 + GENOME_PAD_WORDS; }
int pack_genome(const char* d_ascii, uint64_t n, uint64_t* d_packed, cudaStream_t st);
int synth_genome_packed(uint64_t seed, uint64_t n, uint64_t* d_packed, cudaStream_t st);
int unpack_genome(const uint64_t* d_packed, uint64_t n, char* d_ascii, cudaStream_t st);

// ---- suffix array (sa_build.cu) ----
int build_suffix_array(const uint64_t* d_genome, uint64_t n, uint32_t* d_sa, uint32_t* d_isa, cudaStream_t st,
                       int* rounds_out);
// rank lines (common.cuh): d_lines must hold line_sectors(n) * 32 bytes
int build_rank_lines(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, int bases, uint32_t* d_lines,
                     cudaStream_t st);
int invert_permutation(const uint32_t* d_src, uint64_t n, uint32_t* d_dst, cudaStream_t st);
// sufcheck-style validation: counts adjacent pairs that are out of order / undecided within
// max_chars, and positions where isa[sa[r]] != r.
int check_suffix_array(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                       uint32_t max_chars, uint64_t* bad_order, uint64_t* undecided, uint64_t* bad_perm,
                       cudaStream_t st);
// lcp[r] = LCP(suffix sa[r], suffix sa[r+1]) for r < n-1 (the .sa file's second array)
int compute_lcp(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, uint32_t* d_lcp, cudaStream_t st);
// kflag[r] = 1 iff lcp[r] >= k (r < n-1), kflag[n-1] = 0: all the reference ever asks of the LCP
// array (sa.h:33-57 krmq, sapling_api.h:258,287 countHits)
int compute_kflags(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, int k, uint8_t* d_kflag,
                   cudaStream_t st);

// ---- model (sap_build.cu) ----
struct ModelStats {
  int maxOver, maxUnder, meanError, mostOver, mostUnder;
  uint64_t perfect, nOver, nUnder;
};
// buildPiecewiseLinear + errorStats (sapling_api.h:384-487, 342-379).  d_model must hold
// (1<<nb)+1 entries.  If h_errdump != NULL it receives, per k-mer in text order, (y, predict, val)
// triples as int64 (3*(n-k+1) values) for the errFn dump (:467).
int build_model(const uint64_t* d_genome, uint64_t n, const uint32_t* d_sa, const uint32_t* d_isa,
                const uint8_t* d_kflag, int k, int nb, ModelEntry* d_model, ModelStats* stats,
                int64_t* h_errdump, cudaStream_t st);

// narrow device layout of the model (model.cu)
int build_narrow_model(const ModelEntry* d_model, int nb, int shift, uint2* d_narrow, int* ok, cudaStream_t st);
// checkpoints first .. first+count-1 (of (1<<nb)+1) rebuilt from the narrow table into two int64 device arrays
int widen_model(const uint2* d_narrow, int nb, int shift, long long last_x, long long last_y, uint64_t first,
                uint64_t count, long long* d_xs, long long* d_ys, cudaStream_t st);

// countHitsLeft/Right (sapling_api.h:254-263, 283-289) over kflag
int count_hits(const uint8_t* d_kflag, uint64_t n, int k, const uint32_t* d_sa_pos, size_t count,
               uint32_t maxHits, uint32_t* d_left, uint32_t* d_right, cudaStream_t st);

}  // namespace sb

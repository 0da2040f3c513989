// partition.cuh -- partitioned batch queries (partition.cu): the batch is bucketed by the top bits of the k-mer so that
// the query kernel walks the index slice by slice.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb {

constexpr uint32_t kPartChunk = 16384;  // queries per partition chunk: slots fit 16 bits, a chunk is staged in shared memory
                                        // (8192 with two blocks per SM was measured: no faster, and twice the count table)
constexpr int kPartMaxBits = 11;        // at most 2048 slices
// (kSlotShift, kSlotKmerMask, slot_in_kmer_tag: common.cuh)

// bytes of device scratch launch_partitioned_query needs for nq queries and 2^pbits slices
size_t partition_workspace_bytes(size_t nq, int pbits);
// partition -> query -> un-permute, all enqueued on st.  nq < 2^32.  out[i] = the reference's plQuery answer for kmers[i].
// ev != nullptr: five events recorded on st around the stages (before the histogram, after the scans, after the scatter,
// after the query kernel, after the un-permute) -- sapling_b200_stage_ms.
int launch_partitioned_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, void* ws,
                             int pbits, cudaStream_t st, cudaEvent_t* ev = nullptr);

// query.cu
// d_tiles != nullptr: a zeroed device counter; the kernel then walks the batch in order (query.cu QueryCursor)
int launch_kmer_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, cudaStream_t st,
                      const char** name_out, const uint16_t* d_slot, unsigned long long* d_tiles);

}  // namespace sb

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
// partition -> query -> un-permute, all enqueued on st.  nq < 2^32.  out[i] = the reference's plQuery answer for kmers[i]
// as long long (d_out), or as uint32_t with 0xFFFFFFFF for -1 when d_out32 != nullptr.
// ev != nullptr: five events recorded on st around the stages (before the histogram, after the scans, after the scatter,
// after the query kernel, after the un-permute) -- sapling_b200_stage_ms.
int launch_partitioned_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                             void* ws, int pbits, cudaStream_t st, cudaEvent_t* ev = nullptr);

// query.cu
int launch_kmer_query(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                      int occupancy, cudaStream_t st);
// d_tiles: a zeroed device counter the kernel claims its in-order tiles from; d_slot: slot array or slot_in_kmer_tag()
int launch_kmer_query_ordered(const IndexView& ix, const uint64_t* d_part_kmers, size_t nq, long long* d_res,
                              const uint16_t* d_slot, unsigned long long* d_tiles, cudaStream_t st);
// the same kernel over a batch in the caller's order, answers written directly (d_tiles: 8 bytes of device scratch)
int launch_kmer_query_inorder(const IndexView& ix, const uint64_t* d_kmers, size_t nq, long long* d_out, uint32_t* d_out32,
                              unsigned long long* d_tiles, cudaStream_t st);
const char* kmer_query_kernel_name(bool ordered);
int kmer_query_blocks_per_sm(bool ordered, int occupancy);

}  // namespace sb

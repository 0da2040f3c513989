/*
 * sapling_b200.h -- C ABI of libsapling_b200.so: the B200 (sm_100a) implementation of SAPLING's
 * suffix-array query hot path.
 *
 * This is the drop-in boundary.  Plain pointers and sizes only; every entry point cites the
 * reference interface it replaces (file:line into mkirsche/sapling src/).  The C++ shim
 * include/sapling_api.h wraps these into a `struct Sapling` with the reference's surface.
 *
 * Conventions: functions returning int give 0 on success, <0 on error (message in
 * sapling_b200_last_error()).  A handle owns all host and device memory of one index: it is built / ingested on ONE GPU
 * (the current CUDA device at creation) and may then be replicated onto other GPUs of the box (sapling_b200_open_multi /
 * sapling_b200_replicate); the host-pointer batch calls shard over all of them, the device-pointer calls address the
 * handle's own GPU.  Query calls do not mutate the index and may be issued concurrently from several host threads on
 * distinct streams.  There is no CPU fallback: every call fails if no sm_100-class device is usable.
 * Resident per GPU: the 2-bit genome (n/4 bytes), the rank lines (8 n bytes: the suffix array together with the leading
 * bases of every suffix) and the narrow model (8 bytes per bucket) -- 27.7 GB at 3.1 Gbp.
 */
#ifndef SAPLING_B200_H
#define SAPLING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sapling_b200_index sapling_b200_index;

/* flags for the constructors */
#define SAPLING_B200_QUIET 1u        /* do not print the reference's progress lines to stdout */
#define SAPLING_B200_NO_COMPAT 2u    /* 64-bit-safe window arithmetic instead of the reference's
                                        (int)predicted casts (sapling_api.h:209,225).
                                        Identical results for genomes < 2^31 bp. */
#define SAPLING_B200_KEEP_BUILD 4u   /* keep ISA / k-prefix runs on the device after construction
                                        (needed by count_hits / sa_rank) */

/* ---- construction -------------------------------------------------------------------------- */

/* Sapling::Sapling(refFn, saFn, sapFn, numBuckets, maxMem, k, errorFn)   sapling_api.h:492-676.
 * Reads the FASTA (same cleaning rule, :512-548), loads <saFn> if it exists, else builds the
 * suffix array on the GPU and writes <saFn> in the reference format (:565-577,:593-599); loads
 * <sapFn> if it exists, else builds the piecewise-linear model on the GPU (:384-487) and writes
 * it (:656-674).  nb / maxMem / k = -1 select the reference defaults (:500-510,:387-391).
 * err_fn may be NULL/"" (:396-400,:467).  Returns NULL on error. */
sapling_b200_index *sapling_b200_open(const char *ref_fn, const char *sa_fn, const char *sap_fn, int nb,
                                      int maxMem, int k, const char *err_fn, unsigned flags);

/* Same index from memory: genome = n cleaned bases (ASCII A/C/G/T); sa = rank -> position
 * (uint32, may be NULL: built on the GPU).  The model is built on the GPU. */
sapling_b200_index *sapling_b200_create(const char *genome, uint64_t n, const uint32_t *sa, int nb,
                                        int maxMem, int k, unsigned flags);

/* Same, with the model supplied (contents of a .sap file: xlist/ylist of (1<<nb)+1 entries and
 * the five error bounds maxOver,maxUnder,meanError,mostOver,mostUnder; sapling_api.h:639-645). */
sapling_b200_index *sapling_b200_create_with_model(const char *genome, uint64_t n, const uint32_t *sa,
                                                   int k, int nb, const int64_t *xlist,
                                                   const int64_t *ylist, const int *five, unsigned flags);

/* sapling_b200_open, then copies of the resident index on every other GPU whose bit is set in gpu_mask (bit d = CUDA
 * device d; the handle's own device is implied).  The files are parsed and the index is built once; the copies travel
 * device to device (cudaMemcpyPeer: NVLink / NVSwitch between peers).  SURVEY 8b / 8e. */
sapling_b200_index *sapling_b200_open_multi(const char *ref_fn, const char *sa_fn, const char *sap_fn, int nb,
                                            int maxMem, int k, const char *err_fn, unsigned flags, uint64_t gpu_mask);
/* The same replication for an index created any other way.  Devices that already hold a copy are skipped. */
int sapling_b200_replicate(sapling_b200_index *ix, uint64_t gpu_mask);
/* Number of GPUs that hold the index (1 + replicas). */
int sapling_b200_num_devices(const sapling_b200_index *ix);

/* Private index cache (not a reference format): the 2-bit genome, the 32-bit suffix array, the model, the
 * error bounds and the chromosome table of an open index, written so that sapling_b200_open_cache restores the index
 * with three sequential reads instead of cleaning a FASTA and inverting a 16-bytes-per-base .sa file (sapling_api.h:512-611).
 * The restored index answers every entry point exactly like the one that was saved. */
/* With SAPLING_B200_CACHE=1 in the environment sapling_b200_open keeps such a cache as <sapFn>.b200 and opens from it when
 * it is there and was built with the same k (and nb, if one is asked for); the cache is not checked against the FASTA. */
int sapling_b200_save_cache(const sapling_b200_index *ix, const char *path);
sapling_b200_index *sapling_b200_open_cache(const char *path, unsigned flags);

/* Synthetic genome generated on the device: base[i] = "ACGT"[splitmix64(seed+i)>>62]; suffix
 * array and model built on the GPU.  keep_host_genome != 0 also materialises the ASCII genome
 * on the host (needed for sapling_b200_genome()). */
sapling_b200_index *sapling_b200_create_synthetic(uint64_t seed, uint64_t n, int nb, int maxMem, int k,
                                                  int keep_host_genome, unsigned flags);

void sapling_b200_close(sapling_b200_index *ix);

/* ---- introspection ------------------------------------------------------------------------- */

/* Public members of struct Sapling (sapling_api.h:26,29,44,50): any out pointer may be NULL. */
int sapling_b200_info(const sapling_b200_index *ix, uint64_t *n, int *k, int *nb, int *maxOver,
                      int *maxUnder, int *meanError, int *mostOver, int *mostUnder);
/* Sapling::reference (sapling_api.h:20): NUL-terminated cleaned genome, or NULL if not kept. */
const char *sapling_b200_genome(const sapling_b200_index *ix);
/* Sapling::chrEnds (sapling_api.h:59): number of entries / i-th (end position, name). */
size_t sapling_b200_num_chr(const sapling_b200_index *ix);
uint64_t sapling_b200_chr(const sapling_b200_index *ix, size_t i, const char **name);
/* Sapling::perfectPredictions and the over/under counts of the last model build (:47,:53). */
int sapling_b200_build_stats(const sapling_b200_index *ix, uint64_t *perfect, uint64_t *n_over,
                             uint64_t *n_under);
/* Model checkpoints as the reference holds them (sapling_api.h:65): copies (1<<nb)+1 entries. */
int sapling_b200_model(const sapling_b200_index *ix, int64_t *xlist, int64_t *ylist);
/* Sapling::rev (sapling_api.h:41): copies count entries starting at rank `first` to the host (read out of the rank
 * lines; the plain array is not kept resident). */
int sapling_b200_rev(const sapling_b200_index *ix, uint64_t first, uint64_t count, uint32_t *out);
/* Sapling::sa / lsa.inv (sapling_api.h:38; used by align.cpp:287): rank of text position pos.
 * Needs SAPLING_B200_KEEP_BUILD.  Copies count entries starting at position `first`. */
int sapling_b200_sa_rank(const sapling_b200_index *ix, uint64_t first, uint64_t count, uint32_t *out);
/* Writes the index in the reference's on-disk formats. */
int sapling_b200_write_sap(const sapling_b200_index *ix, const char *path);
int sapling_b200_write_sa(const sapling_b200_index *ix, const char *path);
/* sufcheck-style validation of the resident suffix array (the role of libdivsufsort's sufcheck,
 * suffixarray/libdivsufsort/lib/utils.c:161): adjacent suffixes out of order, pairs undecided
 * within max_chars characters, and (with SAPLING_B200_KEEP_BUILD) positions with isa[sa[r]] != r. */
int sapling_b200_check_sa(const sapling_b200_index *ix, uint32_t max_chars, uint64_t *bad_order,
                          uint64_t *undecided, uint64_t *bad_perm);
/* Device memory held by the index, in bytes. */
uint64_t sapling_b200_device_bytes(const sapling_b200_index *ix);
/* Number of kernels launched for k-mer batches through this handle so far (both batch entry points; a partitioned batch
 * is eight launches). */
uint64_t sapling_b200_launch_count(const sapling_b200_index *ix);
/* Name of the CUDA kernel sapling_b200_query_batch(_dev) launches for a small batch on this index, and the resident
 * blocks per SM it is compiled for. */
const char *sapling_b200_query_kernel(const sapling_b200_index *ix, int *blocks_per_sm);
/* The same for a batch of nq queries: a batch large enough to be partitioned (next entry) runs the in-order
 * kernel over the partitioned k-mers instead. */
const char *sapling_b200_query_kernel_for(const sapling_b200_index *ix, size_t nq, int *blocks_per_sm);
/* How many top k-mer bits sapling_b200_query_batch_dev partitions a batch of nq queries by before the query kernel
 * walks it (2^bits slices of the index); 0 = the batch is answered in the caller's order.
 * The answers are the same either way. */
int sapling_b200_query_partition_bits(const sapling_b200_index *ix, size_t nq);
/* Stage timing of sapling_b200_query_batch_dev (measurement only; bench.py's roofline uses it to time the query kernel
 * on its own when the batch is partitioned).  While profiling is on, every call records CUDA events on its launching
 * stream around its stages.  sapling_b200_stage_ms waits for the profiled calls, returns how many there were and the
 * SUM of their stage times in milliseconds: ms[0] histogram + scans, ms[1] scatter, ms[2] the query kernel,
 * ms[3] un-permute (an unpartitioned call only has ms[2]); the list is then cleared. */
int sapling_b200_profile(sapling_b200_index *ix, int on);
int sapling_b200_stage_ms(sapling_b200_index *ix, double ms[4]);

/* ---- hashing (host, no GPU) ---------------------------------------------------------------- */

/* Sapling::kmerize (sapling_api.h:73-78) and kmerizeAdjusted (:83-90). */
int64_t sapling_b200_kmerize(int k, const char *s);
int64_t sapling_b200_kmerize_adjusted(int k, int length, const char *s);

/* ---- queries ------------------------------------------------------------------------------- */

/* queryBatch: out[i] = plQuery(unpack(kmers[i], k), kmers[i], k)   (sapling_api.h:159-248 with the
 * call shape of sapling_example.cpp:137 / align.cpp:279): genome position, or -1.
 * Host pointers; the library streams chunks through the GPU (pinned buffers are copied directly). */
int sapling_b200_query_batch(sapling_b200_index *ix, const uint64_t *kmers, size_t nq, int64_t *out);

/* The same answers in the narrow transfer format (the host path is bound by PCIe bytes: 16 per query above, 10 here for
 * k = 21): kmers = nq little-endian integers of kmer_bytes bytes each (ceil(2k/8) <= kmer_bytes <= 8; 8 = the uint64
 * array of sapling_b200_query_batch), out[i] = the position as uint32 (n < 2^32 always), 0xFFFFFFFF for -1. */
int sapling_b200_query_batch_u32(sapling_b200_index *ix, const void *kmers, int kmer_bytes, size_t nq, uint32_t *out);

/* The densest upload: kmers = a little-endian BIT stream, k-mer i in bits [i * kmer_bits, (i + 1) * kmer_bits) of the
 * buffer (bit j of the stream = bit j % 8 of byte j / 8), 2k <= kmer_bits <= 64; the buffer holds ceil(nq * kmer_bits / 8)
 * bytes.  kmer_bits = 2k uploads nothing but the k-mers: 5.25 + 4 bytes per query at k = 21.  (kmer_bits = 8 * kmer_bytes
 * is sapling_b200_query_batch_u32.) */
int sapling_b200_query_batch_bits(sapling_b200_index *ix, const void *kmers, int kmer_bits, size_t nq, uint32_t *out);

/* Same on device-resident buffers, enqueued on `stream` (a cudaStream_t; NULL = default stream),
 * no copies, no synchronisation. */
int sapling_b200_query_batch_dev(sapling_b200_index *ix, const uint64_t *d_kmers, size_t nq,
                                 int64_t *d_out, void *stream);
int sapling_b200_query_batch_u32_dev(sapling_b200_index *ix, const uint64_t *d_kmers, size_t nq, uint32_t *d_out,
                                     void *stream);

/* long long plQuery(string s, long kmer, size_t length)   sapling_api.h:159.  s holds slen bases
 * (A/C/G/T), length <= slen is the reference's third argument.  A string with any other byte is REJECTED (-2 / error):
 * the reference compares raw bytes (an 'N' never matches and sorts between 'G' and 'T'), which a 2-bit index cannot
 * reproduce; its callers discard such seeds anyway (align.cpp:283-285). */
int64_t sapling_b200_query_str(sapling_b200_index *ix, const char *s, size_t slen, int64_t kmer,
                               size_t length);
/* Batch of strings: query i is s[offsets[i] .. offsets[i]+slens[i]) with kmers[i] and lengths[i]
 * (lengths == NULL means lengths[i] = slens[i]). */
int sapling_b200_query_str_batch(sapling_b200_index *ix, const char *s, const uint64_t *offsets,
                                 const uint32_t *slens, const uint32_t *lengths, const int64_t *kmers,
                                 size_t nq, int64_t *out);

/* queryPiecewiseLinear (sapling_api.h:98-109) for a batch of k-mers (host pointers). */
int sapling_b200_predict_batch(sapling_b200_index *ix, const uint64_t *kmers, size_t nq, uint64_t *out);

/* countHitsLeft / countHitsRight (sapling_api.h:254-263,:283-289) for a batch of ranks.
 * Needs SAPLING_B200_KEEP_BUILD. */
int sapling_b200_count_hits(sapling_b200_index *ix, const uint32_t *sa_pos, size_t count,
                            uint32_t maxHits, uint32_t *left, uint32_t *right);

/* The seed lookups of align.cpp seed_extend (align.cpp:259-300) for a block of reads, both strands, num_seeds seeds per
 * strand at cur_pos = 0, last/(num_seeds-1)*i, ..., last (:271-275).  reads = concatenated ASCII, read r occupies
 * [read_off[r], read_off[r+1]).  Output slot ((r*2 + strand)*num_seeds + i): ref_pos = the hit position that
 * plQuery(query, kmerize(query), k) returned AND whose k bases equal the seed (:279-285), else -1; for hits
 * sa_pos = Sapling::sa[ref_pos] (:287) and left/right = countHitsLeft/Right(sa_pos, max_hits) (:288-289), i.e. the
 * tuple the reference pushes at :291-297 (its first element is left+right+1).  Seeds with a non-ACGT byte and reads
 * shorter than k give -1.  Needs SAPLING_B200_KEEP_BUILD; max_hits <= 255. */
int sapling_b200_seed_batch(sapling_b200_index *ix, const char *reads, const uint64_t *read_off, size_t n_reads,
                            uint32_t num_seeds, uint32_t max_hits, int64_t *ref_pos, uint32_t *sa_pos,
                            uint32_t *left, uint32_t *right);
/* The same tuples in 10 bytes per seed instead of 20 (ref_pos as uint32 with 0xFFFFFFFF for "no hit", the two counts as
 * bytes), blocks of reads pipelined through the GPU (upload, kernel and download of consecutive blocks overlap). */
int sapling_b200_seed_batch_compact(sapling_b200_index *ix, const char *reads, const uint64_t *read_off, size_t n_reads,
                                    uint32_t num_seeds, uint32_t max_hits, uint32_t *ref_pos, uint32_t *sa_pos,
                                    uint8_t *left, uint8_t *right);
/* Device pointers + stream, compact types, no copies, no synchronisation. */
int sapling_b200_seed_batch_dev(sapling_b200_index *ix, const char *d_reads, const uint64_t *d_read_off,
                                size_t n_reads, uint32_t num_seeds, uint32_t max_hits, uint32_t *d_ref_pos,
                                uint32_t *d_sa_pos, uint8_t *d_left, uint8_t *d_right, void *stream);

/* Number of queries so far whose predicted rank was >= n (reference: out-of-bounds read). */
uint64_t sapling_b200_oob_count(sapling_b200_index *ix);

/* ---- measurement helpers ------------------------------------------------------------------- */

/* Present queries sampled on the device: kmer j = genome[pos_j, pos_j+k), pos_j =
 * splitmix64(seed + first + j) mod (n-k); if mut_seed != 0, odd (first+j) get 1-2 substitutions. */
int sapling_b200_sample_queries_dev(sapling_b200_index *ix, uint64_t seed, uint64_t mut_seed,
                                    uint64_t first, size_t nq, uint64_t *d_kmers, void *stream);
/* Counts d_out[i] whose k bases equal the query (the self-check of sapling_example.cpp:144-154)
 * and the number of -1 answers. */
int sapling_b200_verify_dev(sapling_b200_index *ix, const uint64_t *d_kmers, const int64_t *d_out,
                            size_t nq, uint64_t *n_match, uint64_t *n_minus1, void *stream);
/* P of SURVEY 8d measured on the device: the total number of getLcp calls (sapling_api.h:115-120) the REFERENCE's
 * plQuery makes for these queries (the literal probe sequence, none of this library's shortcuts). */
int sapling_b200_count_probes_dev(sapling_b200_index *ix, const uint64_t *d_kmers, size_t nq, uint64_t *n_probes,
                                  void *stream);
/* Random 32-byte-sector gather over `bytes` of scratch HBM: returns achieved GB/s (the
 * "HBM random-sector roofline" denominator). */
int sapling_b200_gather_bench(uint64_t bytes, uint64_t n_loads, int reps, double *gbps);

/* Generalised gather: n_access random accesses of `gran` (32/64/128) contiguous bytes; chain > 1 makes
 * each thread follow a dependent chain of that many accesses.  Returns 1e9 accesses/s. */
int sapling_b200_gather_bench2(uint64_t bytes, uint64_t n_access, int gran, int chain, int blocks_per_sm,
                               int reps, double *gacc_per_s);

const char *sapling_b200_last_error(void);
const char *sapling_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif

/*
 * sapling_oracle.h -- CPU restatement of the SAPLING suffix-array query hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sapling_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_reference.py,
 * tests/golden/make_golden.py) against the unmodified reference header compiled into
 * oracle/_ref/libsapling_ref.so, and against the committed golden vectors those runs produced.
 *
 * All file:line citations are into /root/reference/src/ (mkirsche/sapling @ 4bbe08e).
 */
#ifndef SAPLING_ORACLE_H
#define SAPLING_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct so_index {
  /* cleaned genome, ASCII A/C/G/T, n bytes followed by one NUL (std::string semantics,
     sapling_api.h:20,548) */
  char *ref;
  uint64_t n;
  /* parameters (sapling_api.h:23-35) */
  int k, nb, maxMem;
  /* rev: rank -> text position (classical SA, sapling_api.h:41,609-611);
     inv: text position -> rank (classical ISA, sa.h:26); lcp[r] = LCP(suffix r, suffix r+1) */
  uint32_t *rev, *inv, *lcp;
  /* krmqb[r] = number of consecutive ranks j >= r with lcp[j] >= k (sa.h:33-43) */
  uint32_t *krmqb;
  /* piecewise linear model: (1<<nb)+1 checkpoints (sapling_api.h:65,406-449) */
  int64_t *xlist, *ylist;
  /* global error bounds (sapling_api.h:50,342-379) */
  int maxOver, maxUnder, meanError, mostOver, mostUnder;
  uint64_t perfect, nOver, nUnder;
  /* chromosome ends: cumulative cleaned length -> name (sapling_api.h:59,536-547) */
  uint64_t *chrEndPos;
  char **chrEndName;
  size_t nChr;
} so_index;

/* ---- k-mer hashing (sapling_api.h:73-90) ---- */
int64_t so_kmerize(int k, const char *s);
int64_t so_kmerize_adjusted(int k, int length, const char *s);
/* inverse of so_kmerize: writes k chars + NUL */
void so_unpack_kmer(uint64_t x, int k, char *out);

/* ---- construction ---- */
/* FASTA cleaning rule (sapling_api.h:512-548, util.h:17-20).  Returns malloc'ed NUL-terminated
   genome; fills *n_out.  chr ends are recorded in ix if ix != NULL. */
char *so_read_fasta(const char *path, uint64_t *n_out, so_index *ix);
/* Same rule applied to an in-memory FASTA text. */
char *so_clean_fasta_text(const char *text, size_t len, uint64_t *n_out, so_index *ix);

/* Suffix array (prefix doubling; the SA of a text is unique so any correct builder matches
   sa.h:82-183 / libdivsufsort) + Kasai LCP (sa.h:192-210, suffixarray/addlcp.cpp:19-50). */
int so_build_sa(so_index *ix);
/* .sa file I/O: [u64 n][u64 inv[n]][u64 n-1][u64 lcp[n-1]] (sapling_api.h:565-577,593-599) */
int so_read_sa_file(so_index *ix, const char *path);
int so_write_sa_file(const so_index *ix, const char *path);

/* .sap build (sapling_api.h:384-487, 309-379) -- needs ref, inv, lcp.  nb = -1 -> auto rule
   (sapling_api.h:387-391).  err_fn may be NULL (sapling_api.h:396-400,467). */
int so_build_sap(so_index *ix, int nb, int maxMem, int k, const char *err_fn);
/* .sap file I/O (sapling_api.h:616-645, 656-674) */
int so_read_sap_file(so_index *ix, const char *path);
int so_write_sap_file(const so_index *ix, const char *path);

/* Convenience: the constructor's behaviour (sapling_api.h:492-676): read FASTA, load-or-build
   .sa, load-or-build .sap (writing what it built).  -1 = defaults. */
so_index *so_open(const char *ref_fn, const char *sa_fn, const char *sap_fn, int nb, int maxMem,
                  int k, const char *err_fn);
/* From memory: cleaned genome (ASCII), optional SA (rank->pos, may be NULL => built here). */
so_index *so_from_memory(const char *genome, uint64_t n, const uint32_t *sa, int nb, int maxMem,
                         int k);
/* Query-only index from parts (no LCP / build arrays). */
so_index *so_from_parts(const char *genome, uint64_t n, const uint32_t *sa, int k, int nb,
                        const int64_t *xlist, const int64_t *ylist, const int *five);
void so_close(so_index *ix);

/* ---- the hot path ---- */
/* queryPiecewiseLinear (sapling_api.h:98-109) */
uint64_t so_predict(const so_index *ix, int64_t x);
/* plQuery (sapling_api.h:159-248).  s has slen chars; length is the third argument of the
   reference call.  probes (optional) is incremented once per getLcp call; flags (optional)
   gets SO_FLAG_* bits for inputs on which the reference has undefined behaviour. */
#define SO_FLAG_PRED_OOB 1u  /* predicted >= n: rev[predicted] out of bounds */
#define SO_FLAG_GALLOP_UB 2u /* s.length() > k gallop ran off either end of the suffix array: the
                                reference hangs (:186-195) or reads out of bounds (:231-240) */
int64_t so_plquery(const so_index *ix, const char *s, size_t slen, int64_t kmer, size_t length,
                   uint32_t *probes, uint32_t *flags);
/* Batch of k-mers: out[i] = plQuery(unpack(kmers[i],k), kmers[i], k).  OpenMP over nthreads.
   probes_total / oob_count optional. */
void so_query_batch(const so_index *ix, const uint64_t *kmers, size_t nq, int64_t *out,
                    int nthreads, uint64_t *probes_total, uint64_t *oob_count);
/* Timed variant used by bench.py: strings are built untimed, the plQuery loop is timed
   (sapling_example.cpp:113-118 untimed, :134-140 timed).  Returns seconds. */
double so_query_batch_timed(const so_index *ix, const uint64_t *kmers, size_t nq, int64_t *out,
                            int nthreads);

/* countHitsLeft/Right (sapling_api.h:254-303) */
uint64_t so_count_hits_right(const so_index *ix, uint64_t sa_pos, uint64_t maxHits);
uint64_t so_count_hits_left(const so_index *ix, uint64_t sa_pos, uint64_t maxHits);

/* The seed lookups of align.cpp seed_extend (:259-300) for a block of reads (both strands, num_seeds seeds per strand
   at cur_pos = 0, last/(num_seeds-1)*i, last; :271-275).  reads = concatenated ASCII, read r = [off[r], off[r+1]).
   Output slot ((r*2+strand)*num_seeds + i): ref_pos = the verified hit position (:279-285) or -1; for hits
   sa_pos = inv[ref_pos] (:287, see oracle/ref_harness.cpp on the unfilled `sa` member), left/right =
   countHitsLeft/Right(sa_pos, maxHits) (:288-289).  Needs inv and lcp. */
void so_seed_batch(const so_index *ix, const char *reads, const uint64_t *off, size_t n_reads, size_t num_seeds,
                   size_t maxHits, int64_t *ref_pos, uint32_t *sa_pos, uint32_t *left, uint32_t *right, int nthreads);

/* Independent range oracle: [*lb, *ub) = ranks whose suffix has s as a prefix (the contract of
   libdivsufsort sa_search, suffixarray/libdivsufsort/lib/utils.c:259-326). */
void so_equal_range(const so_index *ix, const char *s, size_t slen, uint64_t *lb, uint64_t *ub);

/* Synthetic data: base[i] = "ACGT"[splitmix64(seed+i)>>62] */
uint64_t so_splitmix64(uint64_t z);
void so_synth_genome(uint64_t seed, uint64_t n, char *out);

#ifdef __cplusplus
}
#endif
#endif

/*
 * ref_harness.cpp -- C-ABI window onto the UNMODIFIED reference implementation.
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no SAPLING logic of its own: it includes the
 * reference header where it lies (/root/reference/src/sapling_api.h, passed with -I by
 * oracle/Makefile) and forwards to `struct Sapling`.  The build product goes to oracle/_ref/
 * (git-ignored, travels to the GPU box).  Used to (1) pin oracle/sapling_oracle.c, (2) generate
 * tests/golden/, (3) serve as the "reference" CPU baseline in bench.py.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "sapling_api.h" /* the reference, unmodified */

namespace {
struct NullBuf : std::streambuf {
  int overflow(int c) override { return c; }
};
inline std::string unpack(uint64_t x, int k) {
  static const char L[4] = {'A', 'C', 'G', 'T'};
  std::string s((size_t)k, 'A');
  for (int i = 0; i < k; i++) s[(size_t)i] = L[(x >> (2 * (k - 1 - i))) & 3u];
  return s;
}
}  // namespace

extern "C" {

/* Sapling::Sapling(...) sapling_api.h:492.  quiet != 0 silences the constructor's cout chatter. */
void *ref_open(const char *ref_fn, const char *sa_fn, const char *sap_fn, int nb, int maxMem, int k,
               const char *err_fn, int quiet) {
  NullBuf nb_;
  std::streambuf *old = nullptr;
  if (quiet) old = std::cout.rdbuf(&nb_);
  Sapling *s = new Sapling(ref_fn, sa_fn, sap_fn, nb, maxMem, k, err_fn ? err_fn : "");
  if (quiet) std::cout.rdbuf(old);
  /* The reference never fclose()s the .sa/.sap files it writes (sapling_api.h:593-599,656-674);
     in its own drivers process exit flushes them.  Flush here so the files are complete while
     this process lives. */
  fflush(NULL);
  return s;
}

/* The same struct filled member by member instead of through the constructor (which needs a FASTA, a 16n-byte .sa
   file and ~29n bytes of RAM, sapling_api.h:559-611: ~90 GB and ~1 h at 3.1 Gbp).  Every member plQuery reads is public
   (sapling_api.h:19-68): reference, n, k, buckets, rev, xlist, ylist and the five error bounds.  The caller streams the
   parts in chunks (ref_parts_*), so peak host memory is the struct itself (~10.4 bytes per base + 16 per bucket);
   plQuery / queryPiecewiseLinear / binarySearch / getLcp are then the untouched reference code. */
void *ref_from_parts(uint64_t n, int k, int nb, const int *five) {
  Sapling *s = new Sapling();
  s->n = (size_t)n;
  s->k = k;
  s->buckets = nb;
  s->maxOver = five[0]; s->maxUnder = five[1]; s->meanError = five[2];
  s->mostOver = five[3]; s->mostUnder = five[4];
  for (int i = 0; i < 256; i++) s->vals[i] = 0; /* sapling_api.h:494-498 */
  s->vals['A'] = 0; s->vals['C'] = 1; s->vals['G'] = 2; s->vals['T'] = 3;
  s->reference.resize((size_t)n);
  s->rev.resize((size_t)n);
  const size_t count = ((size_t)1 << nb) + 1;
  s->xlist = new long long[count];
  s->ylist = new long long[count];
  return s;
}
void ref_parts_genome(void *h, uint64_t first, uint64_t count, const char *bases) {
  std::memcpy(&((Sapling *)h)->reference[(size_t)first], bases, (size_t)count);
}
void ref_parts_rev(void *h, uint64_t first, uint64_t count, const uint32_t *rev32, int nthreads) {
  size_t *dst = ((Sapling *)h)->rev.data() + first;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < (size_t)count; i++) dst[i] = rev32[i];
}
void ref_parts_model(void *h, uint64_t first, uint64_t count, const long long *xs, const long long *ys) {
  Sapling *s = (Sapling *)h;
  std::memcpy(s->xlist + first, xs, (size_t)count * sizeof(long long));
  std::memcpy(s->ylist + first, ys, (size_t)count * sizeof(long long));
}

void ref_close(void *h) { delete (Sapling *)h; } /* the reference leaks xlist/ylist; so do we */

void ref_info(void *h, uint64_t *n, int *k, int *nb, int *five, uint64_t *perfect) {
  Sapling *s = (Sapling *)h;
  if (n) *n = s->n;
  if (k) *k = s->k;
  if (nb) *nb = s->buckets;
  if (five) {
    five[0] = s->maxOver; five[1] = s->maxUnder; five[2] = s->meanError;
    five[3] = s->mostOver; five[4] = s->mostUnder;
  }
  if (perfect) *perfect = s->perfectPredictions;
}

const char *ref_genome(void *h) { return ((Sapling *)h)->reference.c_str(); }
const long long *ref_xlist(void *h) { return ((Sapling *)h)->xlist; }
const long long *ref_ylist(void *h) { return ((Sapling *)h)->ylist; }
const size_t *ref_rev(void *h) { return ((Sapling *)h)->rev.data(); }
const size_t *ref_inv(void *h) { return ((Sapling *)h)->lsa.inv.data(); }
const size_t *ref_lcp(void *h) { return ((Sapling *)h)->lsa.lcp.data(); }

size_t ref_num_chr(void *h) { return ((Sapling *)h)->chrEnds.size(); }
size_t ref_chr(void *h, size_t i, char *name_out, size_t cap) {
  Sapling *s = (Sapling *)h;
  auto it = s->chrEnds.begin();
  std::advance(it, (long)i);
  std::strncpy(name_out, it->second.c_str(), cap);
  if (cap) name_out[cap - 1] = 0;
  return it->first;
}

long long ref_kmerize(void *h, const char *s) { return ((Sapling *)h)->kmerize(std::string(s)); }
long long ref_kmerize_adjusted(void *h, int length, const char *s) {
  return ((Sapling *)h)->kmerizeAdjusted(length, std::string(s));
}
size_t ref_predict(void *h, long long x) { return ((Sapling *)h)->queryPiecewiseLinear(x); }

/* A predicted rank >= n makes plQuery read rev[] out of bounds (sapling_api.h:162) and go on with whatever it
   finds there -- undefined, and now and then a segfault.  The harness does not run the reference into it: such a query is
   answered REF_UNDEFINED (LLONG_MIN) here, -2 in ref_seed_batch, and the callers leave it out of their comparisons
   (the product clamps the prediction to n - 1 and counts the event). */
static const long long REF_UNDEFINED = (long long)(1ull << 63);
static inline bool ref_prediction_out_of_range(Sapling *s, long long kmer) {
  return s->queryPiecewiseLinear(kmer) >= s->n;
}

/* plQuery(string s, long kmer, size_t length)  sapling_api.h:159 */
long long ref_query_str(void *h, const char *s, size_t slen, long long kmer, size_t length) {
  if (ref_prediction_out_of_range((Sapling *)h, kmer)) return REF_UNDEFINED;
  return ((Sapling *)h)->plQuery(std::string(s, slen), (long)kmer, length);
}

/* out[i] = plQuery(unpack(kmers[i]), kmers[i], k), the call shape of sapling_example.cpp:137.
   Strings are built before the clock starts (sapling_example.cpp:113-118); returns seconds spent
   in the query loop (:134-140). */
double ref_query_batch(void *h, const uint64_t *kmers, size_t nq, long long *out, int nthreads) {
  Sapling *s = (Sapling *)h;
  const int k = s->k;
  std::vector<std::string> queries(nq);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < nq; i++) {
    queries[i] = unpack(kmers[i], k);
    out[i] = ref_prediction_out_of_range(s, (long long)kmers[i]) ? REF_UNDEFINED : 0; /* checked outside the timed loop */
  }
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < nq; i++)
    if (out[i] != REF_UNDEFINED) out[i] = s->plQuery(queries[i].substr(0, (size_t)k), (long)kmers[i], queries[i].length());
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

size_t ref_count_hits_left(void *h, size_t sa_pos, size_t maxHits) {
  return ((Sapling *)h)->countHitsLeft(sa_pos, maxHits);
}
size_t ref_count_hits_right(void *h, size_t sa_pos, size_t maxHits) {
  return ((Sapling *)h)->countHitsRight(sa_pos, maxHits);
}

/* The seed lookups of align.cpp seed_extend (:267-300) for a block of reads, written against the reference's own
   methods.  Output slot ((r*2+strand)*num_seeds + i): ref_pos = verified hit position or -1 (plQuery returned -1, or
   the k bases at the returned position differ from the seed, :280-285; -2: undefined in the reference, predicted rank
   >= n); for hits sa_pos/left/right as pushed at
   :287-297.  The reference reads `sapling->sa[ref_pos]` (:287), a member that is declared (sapling_api.h:38) but never
   filled, so the shipped align crashes there; the intended array is the inverse suffix array lsa.inv (the one
   countHitsLeft/Right index by rank), which is what is used here.  Reads shorter than k are skipped (the reference
   underflows `last`, :261). */
static char ref_complement(char c) { /* align.cpp:241-248 */
  if (c == 'A') return 'T';
  if (c == 'C') return 'G';
  if (c == 'G') return 'C';
  if (c == 'T') return 'A';
  return c;
}
void ref_seed_batch(void *h, const char *reads, const uint64_t *off, size_t n_reads, size_t num_seeds, size_t maxHits,
                    long long *ref_pos_out, uint32_t *sa_pos_out, uint32_t *left_out, uint32_t *right_out, int nthreads) {
  Sapling *sapling = (Sapling *)h;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
  for (size_t r = 0; r < n_reads; r++) {
    std::string readSeq(reads + off[r], (size_t)(off[r + 1] - off[r]));
    for (int iter = 0; iter < 2; iter++) {
      for (size_t i = 0; i < num_seeds; i++) {
        size_t slot = (r * 2 + (size_t)iter) * num_seeds + i;
        ref_pos_out[slot] = -1;
        sa_pos_out[slot] = left_out[slot] = right_out[slot] = 0;
      }
      if (readSeq.length() < (size_t)sapling->k) continue;
      size_t last = readSeq.length() - sapling->k; /* :261 */
      std::string seq = readSeq;
      if (iter) { /* revComp, :251-256 */
        for (size_t j = 0; j < readSeq.length(); j++) seq[j] = ref_complement(readSeq[readSeq.length() - 1 - j]);
      }
      for (size_t i = 0; i < num_seeds; i++) {
        size_t cur_pos = 0; /* :273-275 */
        if (i == num_seeds - 1) cur_pos = last;
        else if (i > 0) cur_pos = last / (num_seeds - 1) * i;
        std::string query = seq.substr(cur_pos, sapling->k);
        long long val = sapling->kmerize(query);
        if (ref_prediction_out_of_range(sapling, val)) { /* undefined in the reference (see REF_UNDEFINED) */
          ref_pos_out[(r * 2 + (size_t)iter) * num_seeds + i] = -2;
          continue;
        }
        long long ref_pos_signed = sapling->plQuery(query, val, sapling->k);
        if (ref_pos_signed == -1) continue;
        size_t ref_pos = (size_t)ref_pos_signed;
        std::string ref_seq = sapling->reference.substr(ref_pos, sapling->k);
        if (query.compare(ref_seq)) continue;
        size_t slot = (r * 2 + (size_t)iter) * num_seeds + i;
        size_t sa_pos = sapling->lsa.inv[ref_pos];
        ref_pos_out[slot] = ref_pos_signed;
        sa_pos_out[slot] = (uint32_t)sa_pos;
        left_out[slot] = (uint32_t)sapling->countHitsLeft(sa_pos, maxHits);
        right_out[slot] = (uint32_t)sapling->countHitsRight(sa_pos, maxHits);
      }
    }
  }
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} /* extern "C" */

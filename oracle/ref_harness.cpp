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

/* plQuery(string s, long kmer, size_t length)  sapling_api.h:159 */
long long ref_query_str(void *h, const char *s, size_t slen, long long kmer, size_t length) {
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
  for (size_t i = 0; i < nq; i++) queries[i] = unpack(kmers[i], k);
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < nq; i++)
    out[i] = s->plQuery(queries[i].substr(0, (size_t)k), (long)kmers[i], queries[i].length());
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

size_t ref_count_hits_left(void *h, size_t sa_pos, size_t maxHits) {
  return ((Sapling *)h)->countHitsLeft(sa_pos, maxHits);
}
size_t ref_count_hits_right(void *h, size_t sa_pos, size_t maxHits) {
  return ((Sapling *)h)->countHitsRight(sa_pos, maxHits);
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} /* extern "C" */

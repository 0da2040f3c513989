/*
 * ref_divsuf_harness.c -- C-ABI window onto the reference's libdivsufsort submodule (mkirsche fork, v2.0.2-1), compiled
 * from /root/reference/suffixarray/libdivsufsort where it lies (oracle/Makefile: divsuf).
 *
 * TEST INFRASTRUCTURE ONLY.  No logic of its own: sa_search (lib/utils.c:259-326) is the independent match-range oracle of
 * SURVEY 8c -- [left, left + count) = the ranks whose suffixes start with the pattern -- and divsufsort
 * (lib/divsufsort.c:332) the suffix-array builder of suffixarray/refToSuffixArray.sh.  Index type: int64 (the reference's
 * own sed, refToSuffixArray.sh:12).
 */
#include <stdint.h>
#include <stdlib.h>

#include "divsufsort.h"

/* count of suffixes that start with P; *left = the first such rank (or the insertion point when count == 0) */
int64_t ref_sa_search(const uint8_t *T, int64_t n, const uint8_t *P, int64_t m, const int64_t *SA, int64_t *left) {
  saidx_t l = -1;
  saidx_t c = sa_search(T, (saidx_t)n, P, (saidx_t)m, (const saidx_t *)SA, (saidx_t)n, &l);
  if (left) *left = (int64_t)l;
  return (int64_t)c;
}

/* many patterns of one length against one index (patterns concatenated) */
void ref_sa_search_batch(const uint8_t *T, int64_t n, const uint8_t *P, int64_t m, int64_t count, const int64_t *SA,
                         int64_t *left, int64_t *cnt) {
  for (int64_t i = 0; i < count; i++) cnt[i] = ref_sa_search(T, n, P + i * m, m, SA, left + i);
}

/* SA[0..n) = the suffix array of T */
int ref_divsufsort(const uint8_t *T, int64_t *SA, int64_t n) { return (int)divsufsort(T, (saidx_t *)SA, (saidx_t)n); }

int ref_sufcheck(const uint8_t *T, const int64_t *SA, int64_t n) { return (int)sufcheck(T, (const saidx_t *)SA, (saidx_t)n, 0); }

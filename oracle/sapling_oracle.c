/*
 * sapling_oracle.c -- CPU restatement (plain C) of the SAPLING query hot path and .sap build.
 *
 * TEST INFRASTRUCTURE ONLY (see sapling_oracle.h).  Parity status: PINNED against the
 * unmodified reference header (oracle/_ref/libsapling_ref.so) and tests/golden/.
 *
 * File:line citations are into /root/reference/src/ (mkirsche/sapling @ 4bbe08e).
 * Compile with -ffp-contract=off: the reference binary (g++ -O2, baseline x86-64) evaluates the
 * interpolation with separate IEEE double multiply / divide / add.
 */
#define _GNU_SOURCE
#include "sapling_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* k-mer hashing                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* vals[] table of sapling_api.h:494-498: A0 C1 G2 T3, every other byte 0 */
static inline int base_code(unsigned char c) {
  switch (c) {
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 0;
  }
}

/* sapling_api.h:73-78 */
int64_t so_kmerize(int k, const char *s) {
  int64_t h = 0;
  for (int i = 0; i < k; i++) h = (int64_t)(((uint64_t)h << 2) | (uint64_t)base_code((unsigned char)s[i]));
  return h;
}

/* sapling_api.h:83-90: shorter strings get a 'G' appended and are left-aligned to 2k bits */
int64_t so_kmerize_adjusted(int k, int length, const char *s) {
  if (length >= k) return so_kmerize(k, s);
  int64_t h = 0;
  for (int i = 0; i < length; i++) h = (int64_t)(((uint64_t)h << 2) | (uint64_t)base_code((unsigned char)s[i]));
  h = (int64_t)(((uint64_t)h << 2) | 2u);
  return (int64_t)((uint64_t)h << (2 * (k - length - 1)));
}

void so_unpack_kmer(uint64_t x, int k, char *out) {
  static const char L[4] = {'A', 'C', 'G', 'T'};
  for (int i = 0; i < k; i++) out[i] = L[(x >> (2 * (k - 1 - i))) & 3u];
  out[k] = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* FASTA cleaning                                                                              */
/* ------------------------------------------------------------------------------------------ */

static void chr_end_set(so_index *ix, uint64_t pos, const char *name, size_t name_len) {
  /* std::map<size_t,string> assignment (sapling_api.h:538,546): same key overwrites */
  for (size_t i = 0; i < ix->nChr; i++) {
    if (ix->chrEndPos[i] == pos) {
      free(ix->chrEndName[i]);
      ix->chrEndName[i] = strndup(name, name_len);
      return;
    }
  }
  ix->chrEndPos = (uint64_t *)realloc(ix->chrEndPos, (ix->nChr + 1) * sizeof(uint64_t));
  ix->chrEndName = (char **)realloc(ix->chrEndName, (ix->nChr + 1) * sizeof(char *));
  ix->chrEndPos[ix->nChr] = pos;
  ix->chrEndName[ix->nChr] = strndup(name, name_len);
  ix->nChr++;
}

/* sapling_api.h:520-548 with getline() semantics: split on '\n'; a line whose first byte is
   '>' is a header (name = text after '>' up to the first blank); every other line contributes
   its bytes after a-z -> A-Z, keeping only A/C/G/T (util.h:17-20). */
char *so_clean_fasta_text(const char *text, size_t len, uint64_t *n_out, so_index *ix) {
  char *out = (char *)malloc(len + 1);
  uint64_t cnt = 0;
  const char *cur_name = NULL;
  size_t cur_name_len = 0;
  size_t p = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && text[e] != '\n') e++;
    /* line = text[p,e) */
    if (e > p && text[p] == '>') {
      if (cur_name_len > 0 && ix) chr_end_set(ix, cnt, cur_name, cur_name_len);
      size_t t = p + 1;
      while (t < e && text[t] != ' ') t++;
      cur_name = text + p + 1;
      cur_name_len = t - (p + 1);
    } else {
      for (size_t i = p; i < e; i++) {
        char c = text[i];
        if (c >= 'a' && c <= 'z') c = (char)(c + 'A' - 'a');
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') out[cnt++] = c;
      }
    }
    p = e + 1;
  }
  if (cur_name_len > 0 && ix) chr_end_set(ix, cnt, cur_name, cur_name_len);
  out[cnt] = 0;
  *n_out = cnt;
  return out;
}

char *so_read_fasta(const char *path, uint64_t *n_out, so_index *ix) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 1);
  size_t got = fread(buf, 1, (size_t)sz, f);
  fclose(f);
  char *g = so_clean_fasta_text(buf, got, n_out, ix);
  free(buf);
  return g;
}

/* ------------------------------------------------------------------------------------------ */
/* Suffix array + LCP                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* LSD radix sort of (key,val) by 64-bit key, 16-bit digits; only the digits below key_bits. */
static void radix_sort_pairs(uint64_t *key, uint32_t *val, uint64_t *key2, uint32_t *val2, size_t n,
                             int key_bits) {
  size_t *cnt = (size_t *)malloc(65536 * sizeof(size_t));
  for (int shift = 0; shift < key_bits; shift += 16) {
    memset(cnt, 0, 65536 * sizeof(size_t));
    for (size_t i = 0; i < n; i++) cnt[(key[i] >> shift) & 0xFFFF]++;
    size_t sum = 0;
    for (size_t d = 0; d < 65536; d++) {
      size_t c = cnt[d];
      cnt[d] = sum;
      sum += c;
    }
    for (size_t i = 0; i < n; i++) {
      size_t d = (key[i] >> shift) & 0xFFFF;
      key2[cnt[d]] = key[i];
      val2[cnt[d]] = val[i];
      cnt[d]++;
    }
    uint64_t *tk = key; key = key2; key2 = tk;
    uint32_t *tv = val; val = val2; val2 = tv;
  }
  free(cnt);
  /* number of passes = ceil(key_bits/16); if odd the result sits in the scratch buffers */
  int passes = (key_bits + 15) / 16;
  if (passes & 1) {
    memcpy(key2, key, n * sizeof(uint64_t));
    memcpy(val2, val, n * sizeof(uint32_t));
  }
}

/* Prefix doubling.  End of text sorts below 'A' (a suffix that is a proper prefix of another is
   the smaller one), which is the order sa.h's DC3 (zero padding, sa.h:18) and libdivsufsort use. */
int so_build_sa(so_index *ix) {
  const uint64_t n = ix->n;
  if (n == 0 || n >= 0xFFFFFFFFull) return -1;
  uint32_t *sa = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint32_t *sa2 = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint32_t *rank = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint64_t *key = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t *key2 = (uint64_t *)malloc(n * sizeof(uint64_t));
  if (!sa || !sa2 || !rank || !key || !key2) return -1;

  /* round 0: order by the first H0 characters, alphabet {$=0,A=1,C=2,G=3,T=4} */
  enum { H0 = 12 };
  for (uint64_t i = 0; i < n; i++) {
    uint64_t v = 0;
    for (int j = 0; j < H0; j++) {
      uint64_t c = (i + j < n) ? (uint64_t)base_code((unsigned char)ix->ref[i + j]) + 1 : 0;
      v = v * 5 + c;
    }
    key[i] = v; /* < 5^12 < 2^28 */
    sa[i] = (uint32_t)i;
  }
  radix_sort_pairs(key, sa, key2, sa2, n, 28);
  uint64_t h = H0;
  for (;;) {
    /* rank = index of the first element of the group of equal keys (1-based) */
    uint64_t groups = 0;
    uint32_t r = 0;
    for (uint64_t i = 0; i < n; i++) {
      if (i == 0 || key[i] != key[i - 1]) {
        r = (uint32_t)(i + 1);
        groups++;
      }
      rank[sa[i]] = r;
    }
    if (groups == n) break;
    int bits = 0;
    while (((uint64_t)1 << bits) <= n + 1) bits++;
    for (uint64_t i = 0; i < n; i++) {
      uint64_t p = sa[i];
      uint64_t r2 = (p + h < n) ? rank[p + h] : 0;
      key[i] = ((uint64_t)rank[p] << bits) | r2;
    }
    radix_sort_pairs(key, sa, key2, sa2, n, 2 * bits);
    h *= 2;
  }
  free(key);
  free(key2);
  free(sa2);

  ix->rev = sa;
  ix->inv = rank;
  for (uint64_t i = 0; i < n; i++) ix->inv[ix->rev[i]] = (uint32_t)i;

  /* Kasai (sa.h:192-210): lcp[r] = LCP(suffix at rank r, suffix at rank r+1) */
  ix->lcp = (uint32_t *)calloc(n > 1 ? n - 1 : 1, sizeof(uint32_t));
  uint64_t curr = 0;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t r = ix->inv[i];
    if (r < n - 1) {
      uint64_t j = ix->rev[r + 1];
      while (i + curr < n && j + curr < n && ix->ref[i + curr] == ix->ref[j + curr]) curr++;
      ix->lcp[r] = (uint32_t)curr;
    }
    if (curr > 0) curr--;
  }
  return 0;
}

static void krmq_init(so_index *ix) {
  /* sa.h:33-43: krmqb has lcp.size()+1 = n entries; krmqb[last] = 0 */
  const uint64_t m = ix->n - 1;
  free(ix->krmqb);
  ix->krmqb = (uint32_t *)malloc((m + 1) * sizeof(uint32_t));
  ix->krmqb[m] = 0;
  for (uint64_t i = m; i-- > 0;)
    ix->krmqb[i] = (ix->lcp[i] < (uint32_t)ix->k) ? 0 : (1 + ix->krmqb[i + 1]);
}

/* sa.h:47-57 */
static inline int query_lcp_k(const so_index *ix, uint64_t a, uint64_t b) {
  uint64_t i = a < b ? a : b;
  uint64_t j = (a < b ? b : a) - 1;
  return (i > j) || ((uint64_t)ix->krmqb[i] + i > j);
}

int so_read_sa_file(so_index *ix, const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  uint64_t sz = 0;
  if (fread(&sz, 8, 1, f) != 1 || sz != ix->n) { fclose(f); return -2; }
  ix->inv = (uint32_t *)malloc(sz * sizeof(uint32_t));
  ix->rev = (uint32_t *)malloc(sz * sizeof(uint32_t));
  enum { CH = 1 << 16 };
  uint64_t *buf = (uint64_t *)malloc(CH * 8);
  for (uint64_t o = 0; o < sz; o += CH) {
    size_t c = (size_t)((sz - o) < CH ? (sz - o) : CH);
    if (fread(buf, 8, c, f) != c) { free(buf); fclose(f); return -3; }
    for (size_t i = 0; i < c; i++) ix->inv[o + i] = (uint32_t)buf[i];
  }
  uint64_t m = 0;
  if (fread(&m, 8, 1, f) != 1) { free(buf); fclose(f); return -4; }
  ix->lcp = (uint32_t *)malloc((m ? m : 1) * sizeof(uint32_t));
  for (uint64_t o = 0; o < m; o += CH) {
    size_t c = (size_t)((m - o) < CH ? (m - o) : CH);
    if (fread(buf, 8, c, f) != c) { free(buf); fclose(f); return -5; }
    for (size_t i = 0; i < c; i++) ix->lcp[o + i] = (uint32_t)buf[i];
  }
  free(buf);
  fclose(f);
  /* sapling_api.h:609-611 */
  for (uint64_t i = 0; i < sz; i++) ix->rev[ix->inv[i]] = (uint32_t)i;
  return 0;
}

int so_write_sa_file(const so_index *ix, const char *path) {
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  uint64_t sz = ix->n;
  fwrite(&sz, 8, 1, f);
  for (uint64_t i = 0; i < sz; i++) { uint64_t v = ix->inv[i]; fwrite(&v, 8, 1, f); }
  uint64_t m = sz - 1;
  fwrite(&m, 8, 1, f);
  for (uint64_t i = 0; i < m; i++) { uint64_t v = ix->lcp[i]; fwrite(&v, 8, 1, f); }
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Model evaluation                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* queryPiecewiseLinear, sapling_api.h:98-109.  The expression
     (long long)(.5 + ylo + (yhi - ylo) * ((x - xlo) * 1. / (xhi - xlo)))
   is evaluated in IEEE double, one rounding per operation, in this order. */
uint64_t so_predict(const so_index *ix, int64_t x) {
  uint64_t bucket = (uint64_t)(x >> (2 * ix->k - ix->nb));
  int64_t xlo = ix->xlist[bucket], xhi = ix->xlist[bucket + 1];
  int64_t ylo = ix->ylist[bucket], yhi = ix->ylist[bucket + 1];
  if (xlo == xhi) return (uint64_t)ylo;
  volatile double num = (double)(x - xlo) * 1.;
  volatile double frac = num / (double)(xhi - xlo);
  volatile double rise = (double)(yhi - ylo) * frac;
  volatile double base = .5 + (double)ylo;
  volatile double sum = base + rise;
  int64_t p = (int64_t)sum;
  if (p < 0) p = 0;
  return (uint64_t)p;
}

/* ------------------------------------------------------------------------------------------ */
/* .sap build                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* getError, sapling_api.h:309-337.  Under-predictions slide y right along the run of ranks that
   share a k-prefix; the over branch computes a shift and then discards it (:325-336). */
static int get_error(const so_index *ix, uint64_t y, uint64_t predict) {
  if (y < predict) {
    int64_t lo = (int64_t)y, hi = (int64_t)predict + 1;
    while (lo < hi - 1) {
      uint64_t mid = (uint64_t)((lo + hi) / 2);
      if (query_lcp_k(ix, y, mid)) lo = (int64_t)mid;
      else hi = (int64_t)mid;
    }
    return (int)(lo - (int64_t)predict);
  }
  if (y == predict) return 0;
  return (int)((int64_t)y - (int64_t)predict);
}

static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

int so_build_sap(so_index *ix, int nb, int maxMem, int k, const char *err_fn) {
  if (k != -1) ix->k = k;           /* sapling_api.h:503-506 */
  if (maxMem != -1) ix->maxMem = maxMem; /* :507-510 */
  ix->nb = nb;
  const uint64_t n = ix->n;
  k = ix->k;
  if (n < (uint64_t)k) return -1;
  /* :387-391 */
  if (ix->nb == -1) {
    ix->nb = 1;
    while ((uint64_t)(1L << ix->nb) * (uint64_t)ix->maxMem * 2 <= n) ix->nb++;
  }
  krmq_init(ix); /* :584,:601 */
  const uint64_t B = (uint64_t)1 << ix->nb;
  const int shift = 2 * k - ix->nb;
  FILE *ef = NULL;
  if (err_fn && err_fn[0]) {
    ef = fopen(err_fn, "w");
    if (ef) fprintf(ef, "%d\n", ix->nb);
  }
  free(ix->xlist);
  free(ix->ylist);
  ix->xlist = (int64_t *)malloc((B + 1) * sizeof(int64_t));
  ix->ylist = (int64_t *)malloc((B + 1) * sizeof(int64_t));
  for (uint64_t i = 0; i <= B; i++) { ix->xlist[i] = -1; ix->ylist[i] = 0; }

  /* pass 1 (:402-434): rolling 2-bit hash in text order */
  const int64_t keep = (int64_t)((1L << (2 * (k - 1))) - 1);
  int64_t hash = so_kmerize(k, ix->ref);
  for (uint64_t i = 0; i + (uint64_t)k <= n; i++) {
    int64_t x = hash;
    hash &= keep;
    hash = (int64_t)((uint64_t)hash << 2);
    if (i + (uint64_t)k < n) hash |= base_code((unsigned char)ix->ref[i + k]);
    uint64_t y = ix->inv[i];
    uint64_t b = (uint64_t)(x >> shift);
    if (ix->xlist[b] == -1 || ix->xlist[b] > x) { ix->xlist[b] = x; ix->ylist[b] = (int64_t)y; }
    if (x > ix->xlist[B]) { ix->xlist[B] = x; ix->ylist[B] = (int64_t)y; }
  }
  /* forward fill (:437-449) */
  if (ix->xlist[0] == -1) { ix->xlist[0] = 0; ix->ylist[0] = 0; }
  for (uint64_t i = 1; i <= B; i++)
    if (ix->xlist[i] == -1) { ix->xlist[i] = ix->xlist[i - 1]; ix->ylist[i] = ix->ylist[i - 1]; }

  /* pass 2 (:451-481): signed error of every k-mer */
  const uint64_t nk = n - (uint64_t)k + 1;
  int *overs = (int *)malloc(nk * sizeof(int));
  int *unders = (int *)malloc(nk * sizeof(int));
  uint64_t no = 0, nu = 0, perfect = 0;
  hash = so_kmerize(k, ix->ref);
  for (uint64_t i = 0; i + (uint64_t)k <= n; i++) {
    uint64_t predict = so_predict(ix, hash);
    uint64_t y = ix->inv[i];
    int val = get_error(ix, y, predict);
    if (ef) fprintf(ef, "%lld %zu %zu %d\n", (long long)hash, (size_t)y, (size_t)predict, val);
    hash &= keep;
    hash = (int64_t)((uint64_t)hash << 2);
    if (i + (uint64_t)k < n) hash |= base_code((unsigned char)ix->ref[i + k]);
    if (val > 0) overs[no++] = val;
    else if (val < 0) unders[nu++] = -val;
    else perfect++;
  }
  if (ef) fclose(ef);

  /* errorStats (:342-379) */
  int maxOver = 0, maxUnder = 0;
  long tot = 0;
  for (uint64_t i = 0; i < no; i++) { if (overs[i] > maxOver) maxOver = overs[i]; tot += labs((long)overs[i]); }
  for (uint64_t i = 0; i < nu; i++) { if (unders[i] > maxUnder) maxUnder = unders[i]; tot += labs((long)unders[i]); }
  if (maxUnder < 2) maxUnder = 2;
  if (maxOver < 2) maxOver = 2;
  uint64_t cnt = no + nu + perfect;
  ix->meanError = (int)(.5 + (double)((unsigned long)tot / cnt));
  qsort(overs, no, sizeof(int), cmp_int);
  qsort(unders, nu, sizeof(int), cmp_int);
  int mostOver = 0, mostUnder = 0;
  const double mostThreshold = 0.95; /* :35 */
  if (no > 0) mostOver = overs[(size_t)(mostThreshold * (double)no)];
  if (nu > 0) mostUnder = unders[(size_t)(mostThreshold * (double)nu)];
  if (mostOver < 1) mostOver = 1;
  if (mostUnder < 1) mostUnder = 1;
  ix->maxOver = maxOver; ix->maxUnder = maxUnder;
  ix->mostOver = mostOver; ix->mostUnder = mostUnder;
  ix->perfect = perfect; ix->nOver = no; ix->nUnder = nu;
  free(overs);
  free(unders);
  return 0;
}

int so_read_sap_file(so_index *ix, const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  int nb = 0;
  if (fread(&nb, sizeof(int), 1, f) != 1) { fclose(f); return -2; }
  uint64_t count = 0;
  if (nb <= 30) { /* :619-628 */
    int c32 = 0;
    if (fread(&c32, sizeof(int), 1, f) != 1) { fclose(f); return -3; }
    count = (uint64_t)(int64_t)c32;
  } else if (fread(&count, 8, 1, f) != 1) { fclose(f); return -3; }
  ix->nb = nb;
  free(ix->xlist); free(ix->ylist);
  ix->xlist = (int64_t *)malloc(count * 8);
  ix->ylist = (int64_t *)malloc(count * 8);
  int five[5];
  if (fread(ix->xlist, 8, count, f) != count || fread(ix->ylist, 8, count, f) != count ||
      fread(five, sizeof(int), 5, f) != 5) { fclose(f); return -4; }
  fclose(f);
  ix->maxOver = five[0]; ix->maxUnder = five[1]; ix->meanError = five[2];
  ix->mostOver = five[3]; ix->mostUnder = five[4];
  return 0;
}

int so_write_sap_file(const so_index *ix, const char *path) {
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  uint64_t count = ((uint64_t)1 << ix->nb) + 1;
  fwrite(&ix->nb, sizeof(int), 1, f);
  if (ix->nb <= 30) { int c32 = (int)count; fwrite(&c32, sizeof(int), 1, f); } /* :659-662 */
  else fwrite(&count, 8, 1, f);
  fwrite(ix->xlist, 8, count, f);
  fwrite(ix->ylist, 8, count, f);
  int five[5] = {ix->maxOver, ix->maxUnder, ix->meanError, ix->mostOver, ix->mostUnder};
  fwrite(five, sizeof(int), 5, f);
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Constructors                                                                                */
/* ------------------------------------------------------------------------------------------ */

static so_index *so_alloc(void) {
  so_index *ix = (so_index *)calloc(1, sizeof(so_index));
  ix->k = 21;      /* sapling_api.h:26 */
  ix->nb = 18;     /* :29 */
  ix->maxMem = 10; /* :32 */
  return ix;
}

static int file_exists(const char *p) {
  FILE *f = p ? fopen(p, "rb") : NULL;
  if (!f) return 0;
  fclose(f);
  return 1;
}

so_index *so_open(const char *ref_fn, const char *sa_fn, const char *sap_fn, int nb, int maxMem,
                  int k, const char *err_fn) {
  so_index *ix = so_alloc();
  if (k != -1) ix->k = k;
  if (maxMem != -1) ix->maxMem = maxMem;
  ix->ref = so_read_fasta(ref_fn, &ix->n, ix);
  if (!ix->ref) { so_close(ix); return NULL; }
  if (file_exists(sa_fn)) {
    if (so_read_sa_file(ix, sa_fn) != 0) { so_close(ix); return NULL; }
  } else {
    if (so_build_sa(ix) != 0) { so_close(ix); return NULL; }
    so_write_sa_file(ix, sa_fn);
  }
  if (file_exists(sap_fn)) {
    if (so_read_sap_file(ix, sap_fn) != 0) { so_close(ix); return NULL; }
  } else {
    if (so_build_sap(ix, nb, maxMem, k, err_fn) != 0) { so_close(ix); return NULL; }
    so_write_sap_file(ix, sap_fn);
  }
  return ix;
}

so_index *so_from_memory(const char *genome, uint64_t n, const uint32_t *sa, int nb, int maxMem,
                         int k) {
  so_index *ix = so_alloc();
  ix->n = n;
  ix->ref = (char *)malloc(n + 1);
  memcpy(ix->ref, genome, n);
  ix->ref[n] = 0;
  if (sa) {
    ix->rev = (uint32_t *)malloc(n * 4);
    ix->inv = (uint32_t *)malloc(n * 4);
    memcpy(ix->rev, sa, n * 4);
    for (uint64_t i = 0; i < n; i++) ix->inv[ix->rev[i]] = (uint32_t)i;
    ix->lcp = (uint32_t *)calloc(n > 1 ? n - 1 : 1, 4);
    uint64_t curr = 0;
    for (uint64_t i = 0; i < n; i++) {
      uint64_t r = ix->inv[i];
      if (r < n - 1) {
        uint64_t j = ix->rev[r + 1];
        while (i + curr < n && j + curr < n && ix->ref[i + curr] == ix->ref[j + curr]) curr++;
        ix->lcp[r] = (uint32_t)curr;
      }
      if (curr > 0) curr--;
    }
  } else if (so_build_sa(ix) != 0) { so_close(ix); return NULL; }
  if (so_build_sap(ix, nb, maxMem, k, NULL) != 0) { so_close(ix); return NULL; }
  return ix;
}

/* Query-only index from parts already in memory (no LCP, no build): used to count probes and to
   answer queries at sizes where rebuilding the model on the CPU would take too long. */
so_index *so_from_parts(const char *genome, uint64_t n, const uint32_t *sa, int k, int nb,
                        const int64_t *xlist, const int64_t *ylist, const int *five) {
  so_index *ix = so_alloc();
  ix->n = n;
  ix->k = k;
  ix->nb = nb;
  ix->ref = (char *)malloc(n + 1);
  memcpy(ix->ref, genome, n);
  ix->ref[n] = 0;
  ix->rev = (uint32_t *)malloc(n * 4);
  memcpy(ix->rev, sa, n * 4);
  uint64_t cnt = ((uint64_t)1 << nb) + 1;
  ix->xlist = (int64_t *)malloc(cnt * 8);
  ix->ylist = (int64_t *)malloc(cnt * 8);
  memcpy(ix->xlist, xlist, cnt * 8);
  memcpy(ix->ylist, ylist, cnt * 8);
  ix->maxOver = five[0]; ix->maxUnder = five[1]; ix->meanError = five[2];
  ix->mostOver = five[3]; ix->mostUnder = five[4];
  return ix;
}

void so_close(so_index *ix) {
  if (!ix) return;
  free(ix->ref); free(ix->rev); free(ix->inv); free(ix->lcp); free(ix->krmqb);
  free(ix->xlist); free(ix->ylist);
  for (size_t i = 0; i < ix->nChr; i++) free(ix->chrEndName[i]);
  free(ix->chrEndPos); free(ix->chrEndName);
  free(ix);
}

/* ------------------------------------------------------------------------------------------ */
/* The hot path                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* getLcp, sapling_api.h:115-120: starts at `start`, trusts the characters before it */
static inline uint64_t get_lcp(const so_index *ix, uint64_t idx, const char *s, uint64_t start,
                               uint64_t length, uint32_t *probes) {
  if (probes) (*probes)++;
  uint64_t i = start;
  while (i < length && idx + i < ix->n && s[i] == ix->ref[idx + i]) i++;
  return i;
}

/* the "suffix is smaller than the query" test used at :143,:167,:175,:214 */
static inline int suffix_is_smaller(const so_index *ix, uint64_t idx, const char *s, uint64_t l) {
  return (l + idx == ix->n) || (s[l] > ix->ref[idx + l]);
}

/* binarySearch, sapling_api.h:133-153 (tail recursion written as a loop) */
static int64_t bounded_search(const so_index *ix, const char *s, uint64_t slen, uint64_t lo,
                              uint64_t hi, uint64_t loLcp, uint64_t hiLcp, uint64_t length,
                              uint32_t *probes) {
  for (;;) {
    if (hi == lo + 2) return (int64_t)(lo + 1); /* :136 -- returned without verification */
    uint64_t mid = (lo + hi) >> 1;
    uint64_t idx = ix->rev[mid];
    uint64_t nLcp = get_lcp(ix, idx, s, loLcp < hiLcp ? loLcp : hiLcp, length, probes);
    if (nLcp == slen) return (int64_t)mid;
    if (lo + 1 >= hi) return -1;
    if (suffix_is_smaller(ix, idx, s, nLcp)) { lo = mid; loLcp = nLcp; }
    else { hi = mid; hiLcp = nLcp; }
  }
}

int64_t so_plquery(const so_index *ix, const char *s, size_t slen, int64_t kmer, size_t length,
                   uint32_t *probes, uint32_t *flags) {
  const uint64_t n = ix->n;
  uint64_t predicted = so_predict(ix, kmer); /* :161 */
  if (predicted >= n) {
    /* reference reads rev[] out of bounds here.  Defined here as: clamp, flag. */
    if (flags) *flags |= SO_FLAG_PRED_OOB;
    predicted = n - 1;
  }
  uint64_t idx = ix->rev[predicted];                      /* :162 */
  uint64_t lcp = get_lcp(ix, idx, s, 0, length, probes);  /* :163 */
  if (lcp == length) return (int64_t)idx;                 /* :164 */
  uint64_t lo, hi, loLcp, hiLcp;
  if (suffix_is_smaller(ix, idx, s, lcp)) {               /* :167 */
    lo = predicted;
    hi = predicted + (uint64_t)(int64_t)ix->mostOver;     /* :171 */
    if (hi > n - 1) hi = n - 1;
    uint64_t hiIdx = ix->rev[hi];
    uint64_t oLcp = get_lcp(ix, hiIdx, s, 0, length, probes);
    if (oLcp == length) return (int64_t)hiIdx;            /* :174 */
    if (suffix_is_smaller(ix, hiIdx, s, oLcp)) {          /* :175 */
      lo = hi;
      loLcp = oLcp;
      hi = predicted + (uint64_t)(int64_t)ix->maxOver + 1; /* :180 */
      if (hi > n - 1) hi = n - 1;
      hiIdx = ix->rev[hi];
      oLcp = get_lcp(ix, hiIdx, s, 0, length, probes);
      if (oLcp == length) return (int64_t)hiIdx;          /* :183 */
      if (slen > (uint64_t)ix->k) {                       /* :184-196 */
        while (oLcp + hiIdx != n && s[oLcp] > ix->ref[hiIdx + oLcp]) {
          /* the reference spins forever once hi is pinned at n-1; defined here as: stop, flag */
          if (hi == n - 1) { if (flags) *flags |= SO_FLAG_GALLOP_UB; break; }
          lo = hi;
          loLcp = oLcp;
          hi += (uint64_t)(int64_t)ix->maxOver;
          if (hi > n - 1) hi = n - 1;
          hiIdx = ix->rev[hi];
          oLcp = get_lcp(ix, hiIdx, s, 0, length, probes);
          if (oLcp == slen) return (int64_t)hiIdx;
        }
      }
      hiLcp = oLcp;
    } else {
      loLcp = lcp;
      hiLcp = oLcp;
    }
  } else {
    /* :209 -- (int) cast of a size_t: wraps for predicted >= 2^31 */
    int32_t p32 = (int32_t)(uint32_t)predicted;
    int32_t v = (int32_t)((uint32_t)p32 - (uint32_t)ix->mostUnder);
    lo = (uint64_t)(int64_t)(v > 0 ? v : 0);
    hi = predicted;
    uint64_t loIdx = ix->rev[lo];
    uint64_t oLcp = get_lcp(ix, loIdx, s, 0, length, probes);
    if (oLcp == slen) return (int64_t)loIdx;              /* :213 */
    if (suffix_is_smaller(ix, loIdx, s, oLcp)) {          /* :214 */
      hiLcp = lcp;
      loLcp = oLcp;
    } else {
      hi = lo;
      hiLcp = oLcp;
      v = (int32_t)((uint32_t)p32 - (uint32_t)ix->maxUnder - 1u); /* :225 */
      lo = (uint64_t)(int64_t)(v > 0 ? v : 0);
      loIdx = ix->rev[lo];
      oLcp = get_lcp(ix, loIdx, s, 0, length, probes);
      if (oLcp == slen) return (int64_t)loIdx;            /* :228 */
      if (slen > (uint64_t)ix->k) {                       /* :229-241 */
        while (oLcp + loIdx != n && s[oLcp] < ix->ref[loIdx + oLcp]) {
          /* the reference does size_t arithmetic here (max((size_t)0,lo) is a no-op), so it reads
             rev[] far out of bounds once lo < maxUnder; defined here as: clamp at 0, stop at 0, flag */
          if (lo == 0) { if (flags) *flags |= SO_FLAG_GALLOP_UB; break; }
          hi = lo;
          hiLcp = oLcp;
          if (lo < (uint64_t)(int64_t)ix->maxUnder) {
            if (flags) *flags |= SO_FLAG_GALLOP_UB;
            lo = 0;
          } else
            lo -= (uint64_t)(int64_t)ix->maxUnder;
          loIdx = ix->rev[lo];
          oLcp = get_lcp(ix, loIdx, s, 0, length, probes);
          if (oLcp == slen) return (int64_t)loIdx;
        }
      }
      loLcp = oLcp;
    }
  }
  int64_t revPos = bounded_search(ix, s, slen, lo, hi, loLcp, hiLcp, length, probes); /* :245 */
  if (revPos == -1) return -1;
  return (int64_t)ix->rev[revPos];                        /* :247 */
}

void so_query_batch(const so_index *ix, const uint64_t *kmers, size_t nq, int64_t *out,
                    int nthreads, uint64_t *probes_total, uint64_t *oob_count) {
  uint64_t ptot = 0, oob = 0;
  const int k = ix->k;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) reduction(+ : ptot, oob) schedule(static)
  for (size_t i = 0; i < nq; i++) {
    char s[40];
    so_unpack_kmer(kmers[i], k, s);
    uint32_t p = 0, fl = 0;
    out[i] = so_plquery(ix, s, (size_t)k, (int64_t)kmers[i], (size_t)k, &p, &fl);
    ptot += p;
    oob += (fl & SO_FLAG_PRED_OOB) ? 1 : 0;
  }
  if (probes_total) *probes_total = ptot;
  if (oob_count) *oob_count = oob;
}

double so_query_batch_timed(const so_index *ix, const uint64_t *kmers, size_t nq, int64_t *out,
                            int nthreads) {
  const int k = ix->k;
  if (nthreads < 1) nthreads = 1;
  char *strs = (char *)malloc(nq * 40);
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < nq; i++) so_unpack_kmer(kmers[i], k, strs + i * 40);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < nq; i++)
    out[i] = so_plquery(ix, strs + i * 40, (size_t)k, (int64_t)kmers[i], (size_t)k, NULL, NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(strs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* sapling_api.h:254-263 */
uint64_t so_count_hits_right(const so_index *ix, uint64_t sa_pos, uint64_t maxHits) {
  for (uint64_t i = 0; i < maxHits; i++)
    if (i + sa_pos > (ix->n - (uint64_t)ix->k) || ix->lcp[i + sa_pos] < (uint32_t)ix->k) return i;
  return maxHits;
}

/* sapling_api.h:283-289 */
uint64_t so_count_hits_left(const so_index *ix, uint64_t sa_pos, uint64_t maxHits) {
  for (uint64_t i = 0; i < maxHits; i++)
    if (sa_pos < i || sa_pos - i >= ix->n - 1 /* lcp has n-1 entries; the reference reads one past for the last rank */ ||
        ix->lcp[sa_pos - i] < (uint32_t)ix->k)
      return i;
  return maxHits;
}

/* align.cpp:241-248 */
static char seed_complement(char c) {
  if (c == 'A') return 'T';
  if (c == 'C') return 'G';
  if (c == 'G') return 'C';
  if (c == 'T') return 'A';
  return c;
}

/* align.cpp:259-300, the seed part of seed_extend.  The comparisons run on the raw read bytes exactly as the
   reference's do (kmerize hashes a non-ACGT byte as 'A', sapling_api.h:494-498, while getLcp and the verify compare see
   the byte itself), so a seed holding an N can never verify. */
void so_seed_batch(const so_index *ix, const char *reads, const uint64_t *off, size_t n_reads, size_t num_seeds,
                   size_t maxHits, int64_t *ref_pos, uint32_t *sa_pos, uint32_t *left, uint32_t *right, int nthreads) {
  const size_t k = (size_t)ix->k;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
  for (size_t r = 0; r < n_reads; r++) {
    const char *read = reads + off[r];
    const size_t len = (size_t)(off[r + 1] - off[r]);
    char *seq = (char *)malloc(len + 1);
    for (int iter = 0; iter < 2; iter++) { /* :267-269 */
      for (size_t i = 0; i < num_seeds; i++) {
        const size_t slot = (r * 2 + (size_t)iter) * num_seeds + i;
        ref_pos[slot] = -1;
        sa_pos[slot] = left[slot] = right[slot] = 0;
      }
      if (len < k) continue; /* the reference underflows `last` here (:261) */
      const size_t last = len - k;
      for (size_t j = 0; j < len; j++) seq[j] = iter ? seed_complement(read[len - 1 - j]) : read[j];
      seq[len] = 0;
      for (size_t i = 0; i < num_seeds; i++) {
        size_t cur_pos = 0; /* :273-275 */
        if (i == num_seeds - 1) cur_pos = last;
        else if (i > 0) cur_pos = last / (num_seeds - 1) * i;
        const char *query = seq + cur_pos;
        const int64_t val = so_kmerize((int)k, query);                          /* :278 */
        const int64_t p = so_plquery(ix, query, k, val, k, NULL, NULL);         /* :279 */
        if (p == -1) continue;                                                  /* :281 */
        if ((uint64_t)p + k > ix->n || memcmp(query, ix->ref + p, k) != 0) continue; /* :283-285 */
        const size_t slot = (r * 2 + (size_t)iter) * num_seeds + i;
        const uint64_t rank = ix->inv[p];                                       /* :287 */
        ref_pos[slot] = p;
        sa_pos[slot] = (uint32_t)rank;
        left[slot] = (uint32_t)so_count_hits_left(ix, rank, maxHits);           /* :288 */
        right[slot] = (uint32_t)so_count_hits_right(ix, rank, maxHits);         /* :289 */
      }
    }
    free(seq);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Independent range oracle                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* three-way compare of s with the suffix at idx: <0 suffix smaller, 0 s is a prefix, >0 bigger */
static int cmp_suffix(const so_index *ix, uint64_t idx, const char *s, size_t slen) {
  for (size_t i = 0; i < slen; i++) {
    if (idx + i >= ix->n) return -1;
    if (ix->ref[idx + i] != s[i]) return ix->ref[idx + i] < s[i] ? -1 : 1;
  }
  return 0;
}

void so_equal_range(const so_index *ix, const char *s, size_t slen, uint64_t *lb, uint64_t *ub) {
  uint64_t lo = 0, hi = ix->n;
  while (lo < hi) {
    uint64_t m = (lo + hi) >> 1;
    if (cmp_suffix(ix, ix->rev[m], s, slen) < 0) lo = m + 1; else hi = m;
  }
  *lb = lo;
  hi = ix->n;
  while (lo < hi) {
    uint64_t m = (lo + hi) >> 1;
    if (cmp_suffix(ix, ix->rev[m], s, slen) <= 0) lo = m + 1; else hi = m;
  }
  *ub = lo;
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic data                                                                              */
/* ------------------------------------------------------------------------------------------ */

uint64_t so_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void so_synth_genome(uint64_t seed, uint64_t n, char *out) {
  static const char L[4] = {'A', 'C', 'G', 'T'};
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) out[i] = L[so_splitmix64(seed + i) >> 62];
}

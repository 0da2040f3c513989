/*
 * sapling_api.h -- drop-in replacement for the reference header of the same name
 * (mkirsche/sapling src/sapling_api.h): `struct Sapling` with the same constructor, public members
 * and methods, backed by libsapling_b200.so (include/sapling_b200.h) instead of host arrays.
 *
 * The reference drivers src/sapling_example.cpp and src/align.cpp compile against this header
 * unchanged (they rely on `using namespace std` and on the <...> headers the reference header
 * pulls in, so this one provides both).  Additions: queryBatch().
 *
 * What is NOT here: any CPU implementation of the query.  Every query runs on the GPU; if the
 * library cannot create the index the constructor prints the error and exits (the reference
 * prints to cerr and continues into undefined behaviour, sapling_api.h:578-581).
 */
#ifndef SAPLING_B200_DROPIN_SAPLING_API_H
#define SAPLING_B200_DROPIN_SAPLING_API_H

#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "sapling_b200.h"

using namespace std; /* the reference gets this from sa.h:11 and its drivers depend on it */

/* util.h:17-20 (the drivers' genome-cleaning predicate) */
inline int bad(char c) { return c != 'A' && c != 'C' && c != 'G' && c != 'T'; }

struct Sapling
{
  /* read-only view of a uint32 device array with the reference's vector<size_t> indexing
     (Sapling::rev sapling_api.h:41, Sapling::sa :38); downloaded on first use */
  struct RankArray
  {
    const Sapling *owner = nullptr;
    int which = 0; /* 0 = rev (rank -> position), 1 = sa (position -> rank) */
    mutable shared_ptr<vector<uint32_t>> host;
    size_t operator[](size_t i) const
    {
      if (!host || host->empty())
      {
        host = make_shared<vector<uint32_t>>(owner->n);
        int rc = which == 0 ? sapling_b200_rev(owner->h.get(), 0, owner->n, host->data())
                            : sapling_b200_sa_rank(owner->h.get(), 0, owner->n, host->data());
        if (rc != 0) { cerr << "sapling_b200: " << sapling_b200_last_error() << endl; exit(1); }
      }
      return (size_t)(*host)[i];
    }
    size_t size() const { return owner ? owner->n : 0; }
  };

  string reference;             /* :20 */
  int alpha = 2;                /* :23 */
  int k = 21;                   /* :26 */
  int buckets = 18;             /* :29 */
  int maxMem = 10;              /* :32 */
  double mostThreshold = 0.95;  /* :35 */
  RankArray sa;                 /* :38 */
  RankArray rev;                /* :41 */
  size_t n = 0;                 /* :44 */
  int maxOver = 0, maxUnder = 0, meanError = 0, mostOver = 0, mostUnder = 0; /* :50 */
  size_t perfectPredictions = 0; /* :53 */
  string errorsFn = "";          /* :56 */
  map<size_t, string> chrEnds;   /* :59 */
  int vals[256];                 /* :62 */

  shared_ptr<sapling_b200_index> h; /* the device-resident index; copies of Sapling share it */

  /* :73-78 */
  long long kmerize(const string &s) { return sapling_b200_kmerize(k, s.c_str()); }
  /* :83-90 */
  long long kmerizeAdjusted(int length, const string &s) { return sapling_b200_kmerize_adjusted(k, length, s.c_str()); }

  /* :98-109 */
  size_t queryPiecewiseLinear(long long x)
  {
    uint64_t in = (uint64_t)x, out = 0;
    check(sapling_b200_predict_batch(h.get(), &in, 1, &out));
    return (size_t)out;
  }

  /* :159-248 -- one query per call (a kernel launch each: use queryBatch for throughput) */
  long long plQuery(string s, long kmer, size_t length)
  {
    long long r = sapling_b200_query_str(h.get(), s.data(), s.length(), (int64_t)kmer, length);
    if (r == -2) { cerr << "sapling_b200: " << sapling_b200_last_error() << endl; exit(1); }
    return r;
  }

  /* addition: out[i] = plQuery(unpack(kmers[i]), kmers[i], k) for the whole batch */
  void queryBatch(const uint64_t *kmers, size_t nq, long long *out)
  {
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    check(sapling_b200_query_batch(h.get(), kmers, nq, reinterpret_cast<int64_t *>(out)));
  }
  vector<long long> queryBatch(const vector<long long> &kmers)
  {
    vector<long long> out(kmers.size());
    queryBatch(reinterpret_cast<const uint64_t *>(kmers.data()), kmers.size(), out.data());
    return out;
  }

  /* addition: the seed lookups of align.cpp:267-300 for a block of reads in one call; slot ((r*2+strand)*num_seeds+i) */
  struct SeedHits
  {
    vector<long long> ref_pos;            /* verified hit position or -1 */
    vector<uint32_t> sa_pos, left, right; /* sa[ref_pos], countHitsLeft/Right(sa_pos, maxHits) */
  };
  SeedHits seedBatch(const vector<string> &reads, size_t numSeeds, size_t maxHits)
  {
    string blob;
    vector<uint64_t> off(reads.size() + 1, 0);
    for (size_t i = 0; i < reads.size(); i++) { blob += reads[i]; off[i + 1] = blob.size(); }
    SeedHits o;
    size_t m = reads.size() * 2 * numSeeds;
    o.ref_pos.resize(m); o.sa_pos.resize(m); o.left.resize(m); o.right.resize(m);
    check(sapling_b200_seed_batch(h.get(), blob.data(), off.data(), reads.size(), (uint32_t)numSeeds, (uint32_t)maxHits,
                                  reinterpret_cast<int64_t *>(o.ref_pos.data()), o.sa_pos.data(), o.left.data(),
                                  o.right.data()));
    return o;
  }

  /* :254-263 */
  size_t countHitsRight(size_t sa_pos, size_t maxHits)
  {
    uint32_t p = (uint32_t)sa_pos, l = 0, r = 0;
    check(sapling_b200_count_hits(h.get(), &p, 1, (uint32_t)maxHits, &l, &r));
    return r;
  }
  /* :283-289 */
  size_t countHitsLeft(size_t sa_pos, size_t maxHits)
  {
    uint32_t p = (uint32_t)sa_pos, l = 0, r = 0;
    check(sapling_b200_count_hits(h.get(), &p, 1, (uint32_t)maxHits, &l, &r));
    return l;
  }

  /* :492 */
  Sapling(string refFnString, string saFnString, string saplingFnString, int numBuckets, int myMaxMem, int myK,
          string errorFn)
  {
    for (int i = 0; i < 256; i++) vals[i] = 0;
    vals['A'] = 0; vals['C'] = 1; vals['G'] = 2; vals['T'] = 3;
    errorsFn = errorFn;
    if (myMaxMem != -1) maxMem = myMaxMem;
    sapling_b200_index *p = sapling_b200_open(refFnString.c_str(), saFnString.c_str(), saplingFnString.c_str(),
                                              numBuckets, myMaxMem, myK, errorFn.c_str(), SAPLING_B200_KEEP_BUILD);
    if (!p) { cerr << "sapling_b200: " << sapling_b200_last_error() << endl; exit(1); }
    h = shared_ptr<sapling_b200_index>(p, sapling_b200_close);
    uint64_t nn = 0;
    sapling_b200_info(p, &nn, &k, &buckets, &maxOver, &maxUnder, &meanError, &mostOver, &mostUnder);
    n = (size_t)nn;
    uint64_t perfect = 0;
    sapling_b200_build_stats(p, &perfect, nullptr, nullptr);
    perfectPredictions = (size_t)perfect;
    const char *g = sapling_b200_genome(p);
    if (g) reference.assign(g, n);
    for (size_t i = 0; i < sapling_b200_num_chr(p); i++)
    {
      const char *name = nullptr;
      uint64_t end = sapling_b200_chr(p, i, &name);
      chrEnds[(size_t)end] = name ? name : "";
    }
    bind();
  }

  Sapling() { for (int i = 0; i < 256; i++) vals[i] = 0; bind(); } /* :678 */
  Sapling(const Sapling &o) { *this = o; }
  Sapling &operator=(const Sapling &o)
  {
    reference = o.reference; alpha = o.alpha; k = o.k; buckets = o.buckets; maxMem = o.maxMem;
    mostThreshold = o.mostThreshold; n = o.n;
    maxOver = o.maxOver; maxUnder = o.maxUnder; meanError = o.meanError; mostOver = o.mostOver; mostUnder = o.mostUnder;
    perfectPredictions = o.perfectPredictions; errorsFn = o.errorsFn; chrEnds = o.chrEnds;
    for (int i = 0; i < 256; i++) vals[i] = o.vals[i];
    h = o.h;
    sa.host = o.sa.host; rev.host = o.rev.host;
    bind();
    return *this;
  }

private:
  void bind() { sa.owner = this; sa.which = 1; rev.owner = this; rev.which = 0; }
  void check(int rc)
  {
    if (rc != 0) { cerr << "sapling_b200: " << sapling_b200_last_error() << endl; exit(1); }
  }
};

#endif

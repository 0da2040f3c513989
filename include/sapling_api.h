/*
 * sapling_api.h -- drop-in replacement for the reference header of the same name
 * (mkirsche/sapling src/sapling_api.h): `struct Sapling` with the same constructor, public members
 * and methods, backed by libsapling_b200.so (include/sapling_b200.h) instead of host arrays.
 *
 * The reference drivers src/sapling_example.cpp and src/align.cpp compile against this header
 * unchanged (they rely on `using namespace std` and on the <...> headers the reference header
 * pulls in, so this one provides both).  Additions: queryBatch().
 *
 * What is NOT here: any CPU implementation of the query.  Every query runs on the GPU.
 * Errors follow the reference's convention -- print to cerr and carry on (sapling_api.h:578-581,:623-649) -- with
 * defined results where the reference runs into undefined behaviour: a constructor that fails leaves an empty index
 * (n = 0), and every query on it, like every failed call, prints the library's message and returns -1 / 0.
 * Set SAPLING_B200_GPUS (a bit mask such as 0xff, or "all") to replicate the index onto several GPUs of the box at
 * construction; queryBatch() then shards every batch over them.
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
        if (rc != 0) cerr << "sapling_b200: " << sapling_b200_last_error() << endl; /* zeros, like a failed read */
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
    if (!check(sapling_b200_predict_batch(h.get(), &in, 1, &out))) return 0;
    return (size_t)out;
  }

  /* :159-248 -- one query per call (a kernel launch each: use queryBatch / plQueryBatch for throughput).
     A string with a byte other than A/C/G/T cannot occur in the cleaned genome (:520-548): the reference returns some
     nearby position that its callers' verification rejects (align.cpp:283-285); here the answer is -1. */
  long long plQuery(string s, long kmer, size_t length)
  {
    for (size_t i = 0; i < s.length(); i++)
      if (bad(s[i])) return -1;
    long long r = sapling_b200_query_str(h.get(), s.data(), s.length(), (int64_t)kmer, length);
    if (r == -2) { cerr << "sapling_b200: " << sapling_b200_last_error() << endl; return -1; }
    return r;
  }

  /* addition: plQuery(queries[i], kmers[i], lengths[i]) for a whole batch of strings in one launch */
  vector<long long> plQueryBatch(const vector<string> &queries, const vector<long long> &kmers, const vector<size_t> &lengths)
  {
    vector<long long> out(queries.size(), -1);
    string blob;
    vector<uint64_t> off;
    vector<uint32_t> slens, lens;
    vector<int64_t> km;
    vector<size_t> which;
    for (size_t i = 0; i < queries.size(); i++)
    {
      bool ok = true;
      for (char c : queries[i]) ok = ok && !bad(c);
      if (!ok) continue; /* -1, see plQuery */
      which.push_back(i);
      off.push_back(blob.size());
      blob += queries[i];
      slens.push_back((uint32_t)queries[i].length());
      lens.push_back((uint32_t)lengths[i]);
      km.push_back((int64_t)kmers[i]);
    }
    vector<int64_t> res(which.size(), -1);
    if (!which.empty() && check(sapling_b200_query_str_batch(h.get(), blob.data(), off.data(), slens.data(), lens.data(),
                                                             km.data(), which.size(), res.data())))
      for (size_t j = 0; j < which.size(); j++) out[which[j]] = res[j];
    return out;
  }

  /* addition: out[i] = plQuery(unpack(kmers[i]), kmers[i], k) for the whole batch (sharded over the GPUs of the index) */
  void queryBatch(const uint64_t *kmers, size_t nq, long long *out)
  {
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    if (!check(sapling_b200_query_batch(h.get(), kmers, nq, reinterpret_cast<int64_t *>(out))))
      for (size_t i = 0; i < nq; i++) out[i] = -1;
  }
  /* the same in the narrow transfer format: kmer_bytes little-endian bytes per k-mer in, uint32 positions out
     (0xFFFFFFFF = -1): 10 bytes per query over PCIe instead of 16 for k = 21 */
  void queryBatchU32(const void *kmers, int kmer_bytes, size_t nq, uint32_t *out)
  {
    if (!check(sapling_b200_query_batch_u32(h.get(), kmers, kmer_bytes, nq, out)))
      for (size_t i = 0; i < nq; i++) out[i] = 0xFFFFFFFFu;
  }
  /* and the densest one: a little-endian bit stream of kmer_bits (2k .. 64) bits per k-mer in (sapling_b200.h) */
  void queryBatchBits(const void *kmers, int kmer_bits, size_t nq, uint32_t *out)
  {
    if (!check(sapling_b200_query_batch_bits(h.get(), kmers, kmer_bits, nq, out)))
      for (size_t i = 0; i < nq; i++) out[i] = 0xFFFFFFFFu;
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
    o.ref_pos.assign(m, -1); o.sa_pos.assign(m, 0); o.left.assign(m, 0); o.right.assign(m, 0);
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
    if (myK != -1) k = myK;
    uint64_t gpus = 0;
    if (const char *e = getenv("SAPLING_B200_GPUS")) gpus = string(e) == "all" ? ~0ull : strtoull(e, nullptr, 0);
    sapling_b200_index *p = sapling_b200_open_multi(refFnString.c_str(), saFnString.c_str(), saplingFnString.c_str(),
                                                    numBuckets, myMaxMem, myK, errorFn.c_str(), SAPLING_B200_KEEP_BUILD, gpus);
    if (!p)
    { /* the reference prints and carries on (:578-581); so does this, with an empty index */
      cerr << "sapling_b200: " << sapling_b200_last_error() << endl;
      bind();
      return;
    }
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
  bool check(int rc)
  {
    if (rc != 0) cerr << "sapling_b200: " << sapling_b200_last_error() << endl;
    return rc == 0;
  }
};

#endif

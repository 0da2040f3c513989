/*
 * align_b200.cpp -- the reference's seed-and-extend aligner (mkirsche/sapling src/align.cpp) with its seed lookup
 * moved to the GPU: BASELINE.json configs[4], "batched GPU seed queries feeding host SSW extension".
 *
 *   align_b200 <reads.fastq> <ref.fa> <out.sam> [num_seeds=7] [sapling_k=16] [flanking_sequence=2] [max_hits=32]
 *              [batch=65536] [threads=<all>]
 *
 * Same command line, same index files (<ref>.sa, <ref>_k<k>.sap) and the same SAM records as the reference's `align`
 * (align.cpp:36-67 flags, :192-225 header, :259-380 seed_extend, :86-146 record format).  What differs is the shape of
 * the work: reads are taken in blocks; ONE call of Sapling::seedBatch (sapling_b200_seed_batch) answers every seed of a
 * block on the GPU -- both strands x num_seeds k-mers per read: plQuery, the verifying compare, sa[ref_pos] and
 * countHitsLeft/Right (align.cpp:277-297) -- and the host keeps what the reference keeps on the host: ordering the
 * seeds by hit count (:301), striped Smith-Waterman extension (:334, the reference's own ssw.c, linked unchanged) and
 * SAM formatting, here spread over all cores with one Aligner per thread.  Records are written in input order.
 *
 * Deliberate differences, both where the reference has undefined behaviour: a read shorter than k is reported
 * unaligned (the reference computes `length - k` in size_t and throws out of substr, align.cpp:261,277), and the
 * reference snapshot never fills Sapling::sa (sapling_api.h:38 / align.cpp:287) -- the intended
 * inverse suffix array is used.
 */
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "sapling_api.h" /* include/sapling_api.h: the drop-in struct Sapling */
#include "ssw_cpp.h"     /* the reference's SSW wrapper, compiled from /root/reference/src as is */

namespace {

struct Options {
  const char *reads = nullptr, *ref = nullptr, *out = nullptr;
  size_t num_seeds = 7, flank = 2, max_hits = 32, batch = 65536;
  int k = 16, threads = 0;
};

bool parse(int argc, char **argv, Options *o) {
  if (argc < 4) return false;
  o->reads = argv[1];
  o->ref = argv[2];
  o->out = argv[3];
  for (int i = 4; i < argc; i++) {
    const char *eq = strchr(argv[i], '=');
    if (!eq) continue;
    const std::string key(argv[i], eq - argv[i]);
    const long v = atol(eq + 1);
    if (key == "num_seeds") o->num_seeds = (size_t)v;
    else if (key == "sapling_k") o->k = (int)v;
    else if (key == "flanking_sequence") o->flank = (size_t)v;
    else if (key == "max_hits") o->max_hits = (size_t)v;
    else if (key == "batch") o->batch = (size_t)(v > 0 ? v : 1);
    else if (key == "threads") o->threads = (int)v;
  }
  return true;
}

struct Read {
  std::string name, seq, qual;
};

/* FASTQ records of four lines; a trailing partial record ends the input (align.cpp:174-190,:236-241) */
size_t read_block(std::ifstream &in, size_t want, std::vector<Read> *block) {
  block->clear();
  std::string l[4];
  while (block->size() < want) {
    int got = 0;
    while (got < 4 && std::getline(in, l[got])) got++;
    if (got < 4) break;
    Read r;
    r.name = l[0].substr(l[0].empty() ? 0 : 1);
    r.seq = l[1];
    r.qual = l[3];
    block->push_back(std::move(r));
  }
  return block->size();
}

std::string reverse_complement(const std::string &s) {
  std::string r(s.size(), 'A');
  for (size_t i = 0; i < s.size(); i++) {
    const char c = s[s.size() - 1 - i];
    r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
  }
  return r;
}

void append(std::string *s, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void append(std::string *s, const char *fmt, ...) {
  char buf[64];
  va_list ap;
  va_start(ap, fmt);
  const int m = vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m > 0) s->append(buf, (size_t)(m < (int)sizeof buf ? m : (int)sizeof buf - 1));
}

/* one SAM line, field for field what align.cpp:86-146 prints */
std::string sam_record(const StripedSmithWaterman::Alignment &a, const Read &rd, const std::string &ref_name, int strand,
                       bool aligned) {
  std::string s = rd.name;
  s += '\t';
  if (!aligned) return s + "4\t*\t0\t255\t*\t*\t0\t0\t*\t*\n";
  /* mapping quality exactly as the reference evaluates it (:100-102), conversions included */
  uint32_t mapq = -4.343 * log(1 - (double)abs(a.sw_score - a.sw_score_next_best) / (double)a.sw_score);
  mapq = (uint32_t)(mapq + 4.99);
  mapq = mapq < 254 ? mapq : 254;
  s += strand ? "16\t" : "0\t";
  s += ref_name;
  append(&s, "\t%d\t%d\t", a.ref_begin + 1, mapq);
  static const char ops[] = "MIDNSHP=X";
  for (size_t c = 0; c < a.cigar.size(); c++) {
    const uint32_t op = a.cigar[c] & 0xfU;
    append(&s, "%lu%c", (unsigned long)(a.cigar[c] >> 4), op > 8 ? 'M' : ops[op]);
  }
  s += "\t*\t0\t0\t";
  s += rd.seq;
  s += '\t';
  if (rd.qual.empty()) s += '*';
  else if (strand) s.append(rd.qual.rbegin(), rd.qual.rend());
  else s += rd.qual;
  append(&s, "\tAS:i:%d", a.sw_score);
  append(&s, "\tNM:i:%d\t", a.mismatches);
  if (a.sw_score_next_best > 0) append(&s, "ZS:i:%d", a.sw_score_next_best);
  s += '\n';
  return s;
}

/* Everything align.cpp:seed_extend does after the seed lookups of one read (:301-379), given those lookups
   (slot (strand * num_seeds + i) of this read in the block's SeedHits). */
std::string extend_read(Sapling &sap, const StripedSmithWaterman::Aligner &ssw, const Options &o, const Read &rd,
                        const Sapling::SeedHits &hits, size_t first_slot) {
  StripedSmithWaterman::Alignment best;
  int best_score = -1, best_strand = 0;
  size_t best_offset = 0;
  bool done = false;
  const size_t k = (size_t)sap.k, len = rd.seq.length();
  if (len >= k) {
    const size_t last = len - k;
    for (int strand = 0; strand < 2 && !done; strand++) {
      const std::string seq = strand ? reverse_complement(rd.seq) : rd.seq;
      /* (hit count, read offset, rank, left, right), ascending: rarest seeds are extended first (:291-301) */
      std::vector<std::tuple<size_t, size_t, size_t, size_t, size_t>> seeds;
      for (size_t i = 0; i < o.num_seeds; i++) {
        const size_t slot = first_slot + (size_t)strand * o.num_seeds + i;
        if (hits.ref_pos[slot] < 0) continue;
        size_t at = 0;
        if (i == o.num_seeds - 1) at = last;
        else if (i > 0) at = last / (o.num_seeds - 1) * i;
        const size_t l = hits.left[slot], r = hits.right[slot];
        seeds.emplace_back(l + r + 1, at, (size_t)hits.sa_pos[slot], l, r);
      }
      std::sort(seeds.begin(), seeds.end());
      for (size_t t = 0; t < seeds.size() && !done; t++) {
        const size_t at = std::get<1>(seeds[t]), rank = std::get<2>(seeds[t]);
        int left = (int)std::get<3>(seeds[t]), right = (int)std::get<4>(seeds[t]);
        if ((size_t)(left + right) > o.max_hits) { /* :311-322 */
          if (best_score == -1) {
            left = std::min(left, (int)(o.max_hits / 2));
            right = std::min(right, (int)(o.max_hits / 2));
          } else {
            left = right = 0;
          }
        }
        for (int d = -left; d <= right && !done; d++) {
          const size_t pos = sap.rev[rank + d];
          long long from = (long long)pos - (long long)at - (long long)o.flank;
          if (from < 0) from = 0;
          const long long to = (long long)pos + (long long)(len - at) + (long long)o.flank;
          if ((size_t)to >= sap.n) continue;
          const int span = (int)(to - from);
          const std::string window = sap.reference.substr((size_t)from, (size_t)span);
          StripedSmithWaterman::Alignment cur;
          StripedSmithWaterman::Filter filter;
          if (!ssw.Align(seq.c_str(), window.c_str(), (int)window.length(), filter, &cur, 15)) continue;
          if ((int)cur.sw_score > best_score) {
            if (cur.mismatches == 0 && cur.cigar.size() == 1) done = true;
            best_score = cur.sw_score;
            best = cur;
            best_offset = (size_t)from;
            best_strand = strand;
          }
        }
      }
    }
  }
  if (best_score < 0) return sam_record(best, rd, "refname", best_strand, false);
  /* chromosome of the hit: smallest end beyond it, and the largest end at or before it (:356-373) */
  const size_t where = (size_t)best.ref_begin + best_offset;
  std::string ref_name = "*";
  size_t end_after = 0, end_before = 0;
  for (const auto &ce : sap.chrEnds) {
    if (ce.first > where && (end_after == 0 || ce.first < end_after)) {
      end_after = ce.first;
      ref_name = ce.second;
    }
    if (ce.first <= where && (end_before == 0 || ce.first > end_before)) end_before = ce.first;
  }
  best.ref_begin += (int32_t)(best_offset - end_before);
  return sam_record(best, rd, ref_name, best_strand, true);
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  if (!parse(argc, argv, &o)) {
    printf("usage: ./align_b200 <query> <ref> <outfile> [num_seeds=<int>] [sapling_k=<int>] [flanking_sequence=<int>] "
           "[max_hits=<int>] [batch=<int>] [threads=<int>]\n");
    return 1;
  }
#ifdef _OPENMP
  if (o.threads > 0) omp_set_num_threads(o.threads);
  const int nthreads = omp_get_max_threads();
#else
  const int nthreads = 1;
#endif
  const std::string ref(o.ref);
  Sapling sap(ref, ref + ".sa", ref + "_k" + std::to_string(o.k) + ".sap", -1, -1, o.k, ""); /* align.cpp:168 */
  std::ifstream in(o.reads);
  FILE *out = fopen(o.out, "w");
  if (!in || !out) {
    fprintf(stderr, "align_b200: cannot open %s\n", !in ? o.reads : o.out);
    return 1;
  }
  cout << "Aligning reads" << endl;
  fprintf(out, "@HD\tVN:1.6\tSO:coordinate\n");
  size_t prev = 0;
  for (const auto &ce : sap.chrEnds) {
    fprintf(out, "@SQ\tSN:%s\tLN:%zu\n", ce.second.c_str(), ce.first - prev);
    prev = ce.first;
  }
  fprintf(out, "@PG\tID:sapling\tVN:1.0\tCL:%s", argv[0]);
  for (int i = 1; i < argc; i++) fprintf(out, " %s", argv[i]);
  fprintf(out, "\n");
  if (sap.n) (void)sap.rev[0]; /* fetch the suffix array to the host once, before the worker threads read it */

  const StripedSmithWaterman::Aligner ssw; /* Align() is const and keeps no state between calls: shared by all threads */
  std::vector<Read> block;
  std::vector<std::string> seqs, records;
  size_t total = 0, seeds = 0;
  double t_seed = 0, t_ext = 0;
  while (read_block(in, o.batch, &block)) {
    seqs.resize(block.size());
    for (size_t i = 0; i < block.size(); i++) seqs[i] = block[i].seq;
    double t0 = 0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#endif
    const Sapling::SeedHits hits = sap.seedBatch(seqs, o.num_seeds, o.max_hits); /* the GPU part */
#ifdef _OPENMP
    t_seed += omp_get_wtime() - t0;
    t0 = omp_get_wtime();
#endif
    records.assign(block.size(), std::string());
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)block.size(); i++) {
      records[(size_t)i] = extend_read(sap, ssw, o, block[(size_t)i], hits, (size_t)i * 2 * o.num_seeds);
    }
#ifdef _OPENMP
    t_ext += omp_get_wtime() - t0;
#endif
    for (const std::string &r : records) fwrite(r.data(), 1, r.size(), out);
    total += block.size();
    seeds += block.size() * 2 * o.num_seeds;
  }
  fclose(out);
  cout << "Aligned " << total << " reads: " << seeds << " seed lookups on the GPU in " << t_seed << " s, extension on "
       << nthreads << " host threads in " << t_ext << " s" << endl;
  return 0;
}

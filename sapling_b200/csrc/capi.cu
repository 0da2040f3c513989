// capi.cu -- host side of libsapling_b200.so: the index object, the loaders for the reference's
// on-disk formats, and the extern "C" surface declared in include/sapling_b200.h.
//
// Mirrors Sapling::Sapling (reference sapling_api.h:492-676): FASTA cleaning, load-or-build of the
// .sa and .sap files, same defaults, same progress lines on stdout (unless SAPLING_B200_QUIET).
// There is no CPU query path in this library: if CUDA is unusable every constructor fails.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sapling_b200.h"
#include "build.cuh"
#include "common.cuh"
#include "partition.cuh"

namespace sb {

// launchers in query.cu
int launch_string_query(const IndexView& ix, const uint64_t* d_words, const uint64_t* d_word_off,
                        const uint32_t* d_slens, const uint32_t* d_lengths, const long long* d_kmers, size_t nq,
                        long long* d_out, cudaStream_t st);
int launch_predict(const IndexView& ix, const uint64_t* d_kmers, size_t nq, uint64_t* d_out, cudaStream_t st);
int launch_seeds(const IndexView& ix, const uint32_t* d_isa, const uint8_t* d_kflag, const char* d_reads,
                 const uint64_t* d_off, size_t n_reads, uint32_t num_seeds, uint32_t maxHits, uint32_t* d_ref_pos,
                 uint32_t* d_sa_pos, uint8_t* d_left, uint8_t* d_right, cudaStream_t st);
int launch_rev_extract(const IndexView& ix, uint64_t first, uint64_t count, uint32_t* d_out, cudaStream_t st);
int launch_sample(const IndexView& ix, uint64_t seed, uint64_t mut_seed, uint64_t first, size_t nq,
                  uint64_t* d_kmers, cudaStream_t st);
int launch_verify(const IndexView& ix, const uint64_t* d_kmers, const long long* d_out, size_t nq,
                  unsigned long long* d_counters, cudaStream_t st);
int launch_probe_count(const IndexView& ix, const uint64_t* d_kmers, size_t nq, unsigned long long* d_total,
                       cudaStream_t st);
int launch_unpack_kmers(const void* d_packed, int kmer_bits, size_t nq, uint64_t* d_kmers, cudaStream_t st);
int run_gather_bench(uint64_t bytes, uint64_t n_loads, int reps, double* gbps);
int run_gather_bench2(uint64_t bytes, uint64_t n_access, int gran, int chain, int blocks_per_sm, int reps,
                      double* gacc_per_s);

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// Experiment settings, read ONCE when an index is created from SAPLING_B200_TUNE="key=value,key=value" (measurement
// scripts and tests only; nothing on the launch path reads the environment).  Defaults are the measured choices.
struct Tuning {
  int partition = 1;                  // part=0: never partition a batch
  size_t partition_min = (size_t)1 << 22;  // part_min=N: smallest batch that is partitioned
  int partition_bits = -1;            // part_bits=B: force the slice count (2^B)
  int occupancy = 0;                  // occ=3..6: resident blocks per SM of kmer_query_kernel (0 = its default; the in-order kernel has one build)
  int hints = -1;                     // hints=<HINT_* bits>
  int line_bases = 0;                 // line_bases=b: leading bases per rank-line entry (tests: forces escapes / ties)
  int chunk_log2 = 0;                 // chunk_log2=l: chunk size of the host-pointer batch path
  int narrow = 1;                     // narrow=0: keep the wide model table
  // inorder_min=N: smallest unpartitioned batch that takes the pipelined in-order kernel (-1: never).  Measured on a 10 Mbp
  // index (tools/inorder_crossover.py, us per batch, in-order / one query per thread): 2^15 27 / 15, 2^18 45 / 31,
  // 2^19 47 / 46, 2^20 76 / 87, 2^22 203 / 303.
  long long inorder_min = 1 << 19;
  static Tuning from_env() {
    Tuning t;
    const char* e = getenv("SAPLING_B200_TUNE");
    if (!e) return t;
    std::string s(e);
    size_t p = 0;
    while (p < s.size()) {
      size_t q = s.find(',', p);
      if (q == std::string::npos) q = s.size();
      const std::string kv = s.substr(p, q - p);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos) {
        const std::string key = kv.substr(0, eq);
        const long long v = atoll(kv.c_str() + eq + 1);
        if (key == "part") t.partition = (int)v;
        else if (key == "part_min") t.partition_min = (size_t)v;
        else if (key == "part_bits") t.partition_bits = (int)v;
        else if (key == "occ") t.occupancy = (int)v;
        else if (key == "hints") t.hints = (int)v;
        else if (key == "line_bases") t.line_bases = (int)v;
        else if (key == "chunk_log2") t.chunk_log2 = (int)v;
        else if (key == "narrow") t.narrow = (int)v;
        else if (key == "inorder_min") t.inorder_min = v;
      }
      p = q + 1;
    }
    return t;
  }
};

}  // namespace sb

using namespace sb;

struct sapling_b200_index {
  int device = 0;
  unsigned flags = 0;
  Tuning tune;
  uint64_t n = 0;
  int k = 21, nb = 18, maxMem = 10;  // sapling_api.h:26,29,32
  ModelStats stats{};
  std::string genome;  // Sapling::reference (may be empty for synthetic indexes)
  std::vector<std::pair<uint64_t, std::string>> chr_ends;  // Sapling::chrEnds, ascending

  uint64_t* d_genome = nullptr;
  uint32_t* d_lines = nullptr;  // rank lines (common.cuh): the resident form of the suffix array
  int line_bases = 0;
  uint32_t* d_sa = nullptr;    // plain rank -> position array: build time, and kept with KEEP_BUILD
  uint32_t* d_isa = nullptr;   // only with KEEP_BUILD
  uint8_t* d_kflag = nullptr;  // only with KEEP_BUILD
  ModelEntry* d_model = nullptr;  // wide checkpoints: build time; resident only when the narrow form cannot hold them
  uint2* d_narrow = nullptr;      // 8-byte-per-bucket layout used by the query kernels (model.cu)
  long long last_x = 0, last_y = 0;
  unsigned hints = 0;
  unsigned long long* d_oob = nullptr;
  uint64_t device_bytes = 0;
  int sa_rounds = 0;

  // replicas of this index on other GPUs (sapling_b200_replicate): the host-pointer batch calls shard over
  // {this, replicas...}; a replica owns its device arrays, shares nothing mutable with the primary
  std::vector<sapling_b200_index*> replicas;
  bool is_replica = false;

  // staging for the host-pointer batch API
  std::mutex mu;
  // Chunks flow through three streams (upload, kernel, download) chained by events, kSlots chunks in flight, so
  // that both copy engines and the SMs stay busy at once: the host-fed rate is then set by PCIe.
  // Chunk size: the first upload and the last download are not overlapped with anything (favours small chunks), but
  // every chunk costs ~40 us of copy-engine / cross-stream latency (favours large ones).  Measured on B200 (c2, 50 M
  // queries): 4 Mi queries per chunk 5.4 G q/s, 2 Mi 5.3, 0.8 Mi 4.5 (profiles/r1u_*); the optimum grows like sqrt(nq).
  static constexpr size_t kChunk = 1u << 22;  // slot capacity, queries
  static constexpr int kSlots = 4;
  std::atomic<uint64_t> launches{0};  // query kernels launched through this handle (sapling_b200_launch_count)
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};  // 0 upload, 1 kernel, 2 download
  cudaEvent_t ev_up[kSlots] = {}, ev_k[kSlots] = {}, ev_down[kSlots] = {};
  uint64_t* d_in[kSlots] = {};    // unpacked k-mers (8 bytes each)
  void* d_raw[kSlots] = {};       // uploaded bytes of the packed-k-mer entry point
  long long* d_out[kSlots] = {};  // answers (8 bytes each, or 4 in the u32 entry point)
  void* h_in[kSlots] = {};
  void* h_out[kSlots] = {};

  // staging of the seed-batch path: three blocks of reads in flight (upload / kernel / download), buffers kept between
  // calls and grown on demand; pinned host mirrors for callers whose own buffers are pageable
  struct SeedSlot {
    char *d_reads = nullptr, *h_reads = nullptr;
    size_t reads_cap = 0;
    uint64_t *d_off = nullptr, *h_off = nullptr;
    size_t off_cap = 0;
    uint32_t *d_rp = nullptr, *d_sp = nullptr, *h_rp = nullptr, *h_sp = nullptr;
    uint8_t *d_l = nullptr, *d_r = nullptr, *h_l = nullptr, *h_r = nullptr;
    size_t seeds_cap = 0;
    void release() {
      cudaFree(d_reads); cudaFree(d_off); cudaFree(d_rp); cudaFree(d_sp); cudaFree(d_l); cudaFree(d_r);
      if (h_reads) cudaFreeHost(h_reads);
      if (h_off) cudaFreeHost(h_off);
      if (h_rp) cudaFreeHost(h_rp);
      if (h_sp) cudaFreeHost(h_sp);
      if (h_l) cudaFreeHost(h_l);
      if (h_r) cudaFreeHost(h_r);
      *this = SeedSlot();
    }
  };
  static constexpr int kSeedSlots = 3;
  SeedSlot seed_slots[kSeedSlots];

  // scratch of the partitioned batch path (partition.cu), one block per stream it was used on, grown on demand
  struct PartWs {
    void* p = nullptr;
    size_t bytes = 0;
  };
  std::mutex mu_ws;
  std::map<cudaStream_t, PartWs> part_ws;

  // stage timing (sapling_b200_profile / sapling_b200_stage_ms): five events per profiled call, see run_kmer_batch
  std::mutex mu_prof;
  bool profiling = false;
  struct ProfCall {
    cudaEvent_t ev[5];
    bool partitioned;
  };
  std::deque<ProfCall> prof_calls;

  // single-query path (plQuery drop-in): one mapped pinned block, no per-call allocation
  std::mutex mu1;
  cudaStream_t s1 = nullptr;
  uint64_t* m1 = nullptr;  // [0..65] packed words, [66] word offset (0), [67] kmer, [68] result, [69] slen|length
  static constexpr uint32_t kSingleMaxBases = 64 * 32;

  IndexView view() const {
    IndexView v;
    v.genome = d_genome;
    v.lines = d_lines;
    v.line_bases = line_bases;
    v.model = d_model;
    v.narrow = d_narrow;
    v.last_x = last_x;
    v.last_y = last_y;
    v.n = n;
    v.k = k;
    v.nb = nb;
    v.shift = 2 * k - nb;
    v.maxOver = stats.maxOver;
    v.maxUnder = stats.maxUnder;
    v.mostOver = stats.mostOver;
    v.mostUnder = stats.mostUnder;
    v.oob_counter = d_oob;
    v.compat = (flags & SAPLING_B200_NO_COMPAT) ? 0 : 1;
    v.hints = hints;
    return v;
  }

  ~sapling_b200_index() {
    for (auto* r : replicas) delete r;
    cudaSetDevice(device);
    for (auto& kv : part_ws) cudaFree(kv.second.p);
    for (auto& c : prof_calls)
      for (int i = 0; i < 5; i++) cudaEventDestroy(c.ev[i]);
    for (int i = 0; i < 3; i++)
      if (streams[i]) cudaStreamDestroy(streams[i]);
    for (int i = 0; i < kSlots; i++) {
      if (ev_up[i]) cudaEventDestroy(ev_up[i]);
      if (ev_k[i]) cudaEventDestroy(ev_k[i]);
      if (ev_down[i]) cudaEventDestroy(ev_down[i]);
      cudaFree(d_in[i]);
      cudaFree(d_raw[i]);
      cudaFree(d_out[i]);
      if (h_in[i]) cudaFreeHost(h_in[i]);
      if (h_out[i]) cudaFreeHost(h_out[i]);
    }
    if (s1) cudaStreamDestroy(s1);
    if (m1) cudaFreeHost(m1);
    for (auto& ss : seed_slots) ss.release();
    cudaFree(d_genome);
    cudaFree(d_lines);
    cudaFree(d_sa);
    cudaFree(d_isa);
    cudaFree(d_kflag);
    cudaFree(d_model);
    cudaFree(d_narrow);
    cudaFree(d_oob);
  }
};

namespace {

struct Say {
  bool on;
  explicit Say(unsigned flags) : on(!(flags & SAPLING_B200_QUIET)) {}
  void operator()(const char* fmt, ...) const {
    if (!on) return;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    fflush(stdout);
  }
};

int check_device(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return -1;
  }
  if (prop.major < 10) {
    set_error("device %d (%s, sm_%d%d) is not a Blackwell sm_100-class GPU; this library is built for sm_100a only",
              dev, prop.name, prop.major, prop.minor);
    return -1;
  }
  return 0;
}

int require_device(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no usable CUDA device: %s (libsapling_b200 has no CPU fallback)", cudaGetErrorString(e));
    return -1;
  }
  if (check_device(dev)) return -1;
  *dev_out = dev;
  return 0;
}

template <typename T>
int dev_alloc(sapling_b200_index* ix, T** p, uint64_t count) {
  const uint64_t bytes = (count ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%llu bytes) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return -1;
  }
  ix->device_bytes += bytes;
  return 0;
}
template <typename T>
void dev_free(sapling_b200_index* ix, T** p, uint64_t count) {
  if (!*p) return;
  cudaFree(*p);
  *p = nullptr;
  ix->device_bytes -= (count ? count : 1) * sizeof(T);
}

// FASTA cleaning rule of sapling_api.h:520-548 / util.h:17-20
void clean_fasta(const char* text, size_t len, std::string* out,
                 std::vector<std::pair<uint64_t, std::string>>* ends) {
  out->clear();
  out->reserve(len);
  std::string cur_name;
  auto set_end = [&](uint64_t pos, const std::string& name) {
    for (auto& e : *ends)
      if (e.first == pos) { e.second = name; return; }  // std::map assignment semantics
    ends->push_back({pos, name});
  };
  size_t p = 0;
  while (p < len) {
    const char* nl = static_cast<const char*>(memchr(text + p, '\n', len - p));
    const size_t e = nl ? (size_t)(nl - text) : len;
    if (e > p && text[p] == '>') {
      if (!cur_name.empty()) set_end(out->size(), cur_name);
      size_t t = p + 1;
      while (t < e && text[t] != ' ') t++;
      cur_name.assign(text + p + 1, t - (p + 1));
    } else {
      for (size_t i = p; i < e; i++) {
        char c = text[i];
        if (c >= 'a' && c <= 'z') c = (char)(c + 'A' - 'a');
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') out->push_back(c);
      }
    }
    p = e + 1;
  }
  if (!cur_name.empty()) set_end(out->size(), cur_name);
}

bool file_exists(const char* p) {
  if (!p || !p[0]) return false;
  FILE* f = fopen(p, "rb");
  if (!f) return false;
  fclose(f);
  return true;
}

// upload ASCII genome and pack it on the device
int upload_genome(sapling_b200_index* ix, const char* genome, uint64_t n) {
  char* d_ascii = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_ascii, n + 64));
  cudaError_t e = cudaMemcpy(d_ascii, genome, n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d_ascii); SB_CUDA_CHECK(e); }
  if (dev_alloc(ix, &ix->d_genome, packed_words(n))) { cudaFree(d_ascii); return -1; }
  int rc = pack_genome(d_ascii, n, ix->d_genome, 0);
  e = cudaDeviceSynchronize();
  cudaFree(d_ascii);
  if (rc) return rc;
  SB_CUDA_CHECK(e);
  return 0;
}

// Pinned double buffer for streaming a file through the GPU: the read of piece i+1 overlaps the copy of piece i.
struct PinnedPair {
  void* h[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  size_t bytes = 0;
  int init(size_t b) {
    bytes = b;
    for (int i = 0; i < 2; i++) {
      SB_CUDA_CHECK(cudaMallocHost(&h[i], b));
      SB_CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    return 0;
  }
  ~PinnedPair() {
    for (int i = 0; i < 2; i++) {
      if (h[i]) cudaFreeHost(h[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
  }
};

// [u64 n][u64 inv[n]][u64 m][u64 lcp[m]]   (sapling_api.h:565-577): only inv is needed.  The 8-byte entries are narrowed
// on the host into a pinned buffer whose upload overlaps the next read.
int read_sa_file(sapling_b200_index* ix, const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("cannot open %s", path); return -1; }
  uint64_t sz = 0;
  if (fread(&sz, 8, 1, f) != 1 || sz != ix->n) {
    fclose(f);
    set_error("%s: suffix array size %llu does not match genome length %llu", path, (unsigned long long)sz,
              (unsigned long long)ix->n);
    return -1;
  }
  if (dev_alloc(ix, &ix->d_isa, sz)) { fclose(f); return -1; }
  const size_t CH = 1u << 22;
  std::vector<uint64_t> buf(CH);
  PinnedPair pp;
  if (pp.init(CH * 4)) { fclose(f); return -1; }
  int slot = 0;
  for (uint64_t o = 0; o < sz; o += CH, slot ^= 1) {
    const size_t c = (size_t)std::min<uint64_t>(CH, sz - o);
    if (fread(buf.data(), 8, c, f) != c) { fclose(f); set_error("Error reading suffix array from file"); return -1; }
    cudaEventSynchronize(pp.ev[slot]);  // the copy that last used this buffer
    uint32_t* h32 = static_cast<uint32_t*>(pp.h[slot]);
    for (size_t i = 0; i < c; i++) h32[i] = (uint32_t)buf[i];
    cudaError_t e = cudaMemcpyAsync(ix->d_isa + o, h32, c * 4, cudaMemcpyHostToDevice, 0);
    if (e != cudaSuccess) { fclose(f); SB_CUDA_CHECK(e); }
    cudaEventRecord(pp.ev[slot], 0);
  }
  fclose(f);
  if (dev_alloc(ix, &ix->d_sa, sz)) return -1;
  // rev[inv[i]] = i  (sapling_api.h:609-611)
  if (invert_permutation(ix->d_isa, sz, ix->d_sa, 0)) return -1;
  SB_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

// The plain rank -> position array of an index whose resident form is the rank lines: ix->d_sa if it is still there,
// else a scratch copy extracted from the lines (freed by the guard).
struct SaGuard {
  const uint32_t* sa = nullptr;
  uint32_t* scratch = nullptr;
  ~SaGuard() { if (scratch) cudaFree(scratch); }
  int get(const sapling_b200_index* ix) {
    if (ix->d_sa) { sa = ix->d_sa; return 0; }
    SB_CUDA_CHECK(cudaMalloc(&scratch, (ix->n ? ix->n : 1) * 4));
    if (launch_rev_extract(ix->view(), 0, ix->n, scratch, 0)) return -1;
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    sa = scratch;
    return 0;
  }
};

int write_sa_file(const sapling_b200_index* ix, const char* path) {
  const uint64_t n = ix->n;
  SaGuard sg;
  if (sg.get(ix)) return -1;
  uint32_t* d_isa_tmp = nullptr;
  const uint32_t* d_isa = ix->d_isa;
  if (!d_isa) {
    SB_CUDA_CHECK(cudaMalloc(&d_isa_tmp, (n ? n : 1) * 4));
    if (invert_permutation(sg.sa, n, d_isa_tmp, 0)) { cudaFree(d_isa_tmp); return -1; }
    d_isa = d_isa_tmp;
  }
  FILE* f = fopen(path, "wb");
  if (!f) { cudaFree(d_isa_tmp); set_error("cannot write %s", path); return -1; }
  uint32_t* d_lcp = nullptr;
  if (cudaMalloc(&d_lcp, (n ? n : 1) * 4) != cudaSuccess) { cudaGetLastError(); cudaFree(d_isa_tmp); fclose(f); set_error("write_sa: allocation failed"); return -1; }
  if (compute_lcp(ix->d_genome, n, sg.sa, d_lcp, 0)) { cudaFree(d_lcp); cudaFree(d_isa_tmp); fclose(f); return -1; }
  const size_t CH = 1u << 22;
  std::vector<uint64_t> buf(CH);
  std::vector<uint32_t> buf32(CH);
  auto dump = [&](const uint32_t* d, uint64_t count) -> int {
    fwrite(&count, 8, 1, f);
    for (uint64_t o = 0; o < count; o += CH) {
      const size_t c = (size_t)std::min<uint64_t>(CH, count - o);
      if (cudaMemcpy(buf32.data(), d + o, c * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
      for (size_t i = 0; i < c; i++) buf[i] = buf32[i];
      if (fwrite(buf.data(), 8, c, f) != c) return -1;
    }
    return 0;
  };
  int rc = dump(d_isa, n);
  if (!rc) rc = dump(d_lcp, n - 1);
  cudaFree(d_lcp);
  cudaFree(d_isa_tmp);
  fclose(f);
  if (rc) set_error("error writing %s", path);
  return rc;
}

// [int nb][int|size_t count][i64 xlist][i64 ylist][5 x int]   (sapling_api.h:616-645)
int read_sap_file(sapling_b200_index* ix, const char* path, std::vector<int64_t>* xs, std::vector<int64_t>* ys) {
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("cannot open %s", path); return -1; }
  int nb = 0;
  uint64_t count = 0;
  bool ok = fread(&nb, sizeof(int), 1, f) == 1;
  if (ok && nb <= 30) {
    int c32 = 0;
    ok = fread(&c32, sizeof(int), 1, f) == 1;
    count = (uint64_t)(int64_t)c32;
  } else if (ok) {
    ok = fread(&count, 8, 1, f) == 1;
  }
  if (!ok || nb < 1 || nb > 31 || count != (1ull << nb) + 1) {
    fclose(f);
    set_error("Error reading sapling data structure from file %s", path);
    return -1;
  }
  xs->resize(count);
  ys->resize(count);
  int five[5];
  ok = fread(xs->data(), 8, count, f) == count && fread(ys->data(), 8, count, f) == count &&
       fread(five, sizeof(int), 5, f) == 5;
  fclose(f);
  if (!ok) { set_error("Error reading sapling data structure from file %s", path); return -1; }
  ix->nb = nb;  // buckets is overwritten from the file (:618)
  ix->stats.maxOver = five[0]; ix->stats.maxUnder = five[1]; ix->stats.meanError = five[2];
  ix->stats.mostOver = five[3]; ix->stats.mostUnder = five[4];
  return 0;
}

int upload_model(sapling_b200_index* ix, const int64_t* xs, const int64_t* ys) {
  const uint64_t count = (1ull << ix->nb) + 1;
  std::vector<ModelEntry> m(count);
  for (uint64_t i = 0; i < count; i++) { m[i].x = xs[i]; m[i].y = ys[i]; }
  if (dev_alloc(ix, &ix->d_model, count)) return -1;
  SB_CUDA_CHECK(cudaMemcpy(ix->d_model, m.data(), count * sizeof(ModelEntry), cudaMemcpyHostToDevice));
  return 0;
}

int validate_params(sapling_b200_index* ix) {
  if (ix->k < 1 || ix->k > 32) { set_error("k=%d out of range (1..32)", ix->k); return -1; }
  if (ix->n < (uint64_t)ix->k + 1) { set_error("genome (%llu bp) shorter than k+1", (unsigned long long)ix->n); return -1; }
  if (ix->n >= 0xFFFFFF00ull) { set_error("genome of %llu bp needs a suffix array wider than 32 bits", (unsigned long long)ix->n); return -1; }
  return 0;
}

// after the model is on the device (wide form): bounds, narrow table, L2 policy
int finish_model(sapling_b200_index* ix) {
  if (ix->nb < 1 || ix->nb > 31 || ix->nb > 2 * ix->k) {
    set_error("nb=%d out of range for k=%d (need 1 <= nb <= min(2k,31))", ix->nb, ix->k);
    return -1;
  }
  const ModelStats& s = ix->stats;
  if (s.maxOver < 0 || s.maxUnder < 0 || s.mostOver < 0 || s.mostUnder < 0) {
    set_error("negative error bounds in model");
    return -1;
  }
  if (!ix->d_oob) {
    if (dev_alloc(ix, &ix->d_oob, 1)) return -1;
    SB_CUDA_CHECK(cudaMemset(ix->d_oob, 0, 8));
  }
  // measured (gpurun r2d / r2e): with the batch partitioned the rank-line slice is what L2 should hold -- evict_last on it
  // makes the c3 kernel time stable where the other policies flip between two regimes
  ix->hints = ix->tune.hints >= 0 ? (unsigned)ix->tune.hints
                                  : (HINT_GENOME_KEEP | HINT_MODEL_KEEP | HINT_SA_KEEP | HINT_IO_STREAM);
  const uint64_t B = 1ull << ix->nb;
  ModelEntry last;
  SB_CUDA_CHECK(cudaMemcpy(&last, ix->d_model + B, sizeof(last), cudaMemcpyDeviceToHost));
  ix->last_x = last.x;
  ix->last_y = last.y;
  const int shift = 2 * ix->k - ix->nb;
  if (!ix->d_narrow && ix->tune.narrow && shift >= 0 && shift <= 31) {
    if (dev_alloc(ix, &ix->d_narrow, B + 1)) return -1;  // + the pad entry (model.cu)
    int ok = 0;
    if (build_narrow_model(ix->d_model, ix->nb, shift, ix->d_narrow, &ok, 0)) return -1;
    if (!ok) dev_free(ix, &ix->d_narrow, B + 1);
  }
  // the narrow table reconstructs xlist / ylist exactly (model.cu widen_model): the 16-byte-per-bucket table goes
  if (ix->d_narrow) dev_free(ix, &ix->d_model, B + 1);
  return 0;
}

// builds whatever is missing: SA (+ISA), rank lines, kflags, model.  model_given: d_model & stats already set.
int build_missing(sapling_b200_index* ix, bool model_given, const char* err_fn, const Say& say) {
  const uint64_t n = ix->n;
  const bool keep = (ix->flags & SAPLING_B200_KEEP_BUILD) != 0;
  if (!ix->d_sa) {
    say("Building suffix array\n");
    if (dev_alloc(ix, &ix->d_sa, n) || dev_alloc(ix, &ix->d_isa, n)) return -1;
    if (build_suffix_array(ix->d_genome, n, ix->d_sa, ix->d_isa, 0, &ix->sa_rounds)) return -1;
    say("Built suffix array of size %llu\n", (unsigned long long)n);
  }
  ix->line_bases = line_bases_for(n);
  if (ix->tune.line_bases >= 4 && ix->tune.line_bases <= kLineMaxBases) ix->line_bases = ix->tune.line_bases;
  if (dev_alloc(ix, &ix->d_lines, line_sectors(n) * 8)) return -1;
  if (build_rank_lines(ix->d_genome, n, ix->d_sa, ix->line_bases, ix->d_lines, 0)) return -1;
  const bool need_build_arrays = !model_given || keep;
  if (need_build_arrays) {
    if (!ix->d_isa) {
      if (dev_alloc(ix, &ix->d_isa, n)) return -1;
      if (invert_permutation(ix->d_sa, n, ix->d_isa, 0)) return -1;
    }
    if (dev_alloc(ix, &ix->d_kflag, n)) return -1;
    if (compute_kflags(ix->d_genome, n, ix->d_sa, ix->k, ix->d_kflag, 0)) return -1;
  }
  if (!model_given) {
    say("Building Sapling\n");
    if (ix->nb == -1) {  // sapling_api.h:387-391
      ix->nb = 1;
      while ((uint64_t)(1ull << ix->nb) * (uint64_t)ix->maxMem * 2 <= n) ix->nb++;
    }
    say("Buckets (log): %d\n", ix->nb);
    if (ix->nb < 1 || ix->nb > 31 || ix->nb > 2 * ix->k) {
      set_error("nb=%d out of range for k=%d (need 1 <= nb <= min(2k,31))", ix->nb, ix->k);
      return -1;
    }
    if (dev_alloc(ix, &ix->d_model, (1ull << ix->nb) + 1)) return -1;
    std::vector<int64_t> dump;
    const bool want_dump = err_fn && err_fn[0];
    const uint64_t nk = n - (uint64_t)ix->k + 1;
    if (want_dump) dump.resize(nk * 3);
    if (build_model(ix->d_genome, n, ix->d_sa, ix->d_isa, ix->d_kflag, ix->k, ix->nb, ix->d_model, &ix->stats,
                    want_dump ? dump.data() : nullptr, 0))
      return -1;
    say("Computing error stats\n");
    say("All overestimates within: %d\n", ix->stats.maxOver);
    say("All underestimates within: %d\n", ix->stats.maxUnder);
    say("Prefect predictions: %llu\n", (unsigned long long)ix->stats.perfect);
    say("Mean error: %d\n", ix->stats.meanError);
    say("0.95 of overestimates within: %d\n", ix->stats.mostOver);
    say("0.95 of underestimates within: %d\n", ix->stats.mostUnder);
    if (want_dump) {  // :396-400,:467 -- "x y predict val" per k-mer, first line nb
      FILE* ef = fopen(err_fn, "w");
      if (ef) {
        fprintf(ef, "%d\n", ix->nb);
        if (!ix->genome.empty()) {
          uint64_t hash = (uint64_t)sapling_b200_kmerize(ix->k, ix->genome.c_str());
          const uint64_t keep_mask = ix->k >= 32 ? ~0ull >> 2 : ((1ull << (2 * (ix->k - 1))) - 1);
          for (uint64_t i = 0; i < nk; i++) {
            fprintf(ef, "%lld %zu %zu %d\n", (long long)hash, (size_t)dump[3 * i], (size_t)dump[3 * i + 1],
                    (int)dump[3 * i + 2]);
            hash = (hash & keep_mask) << 2;
            if (i + ix->k < n) hash |= (uint64_t)sapling_b200_kmerize(1, ix->genome.c_str() + i + ix->k);
          }
        }
        fclose(ef);
      }
    }
  }
  SB_CUDA_CHECK(cudaDeviceSynchronize());
  return finish_model(ix);
}

// what stays resident after construction: the rank lines replace the plain suffix array; the inverse suffix array and
// the lcp >= k flags stay only with KEEP_BUILD (count_hits / sa_rank / seed_batch)
void drop_build_arrays(sapling_b200_index* ix) {
  const bool keep = (ix->flags & SAPLING_B200_KEEP_BUILD) != 0;
  if (!keep) {
    dev_free(ix, &ix->d_sa, ix->n);
    dev_free(ix, &ix->d_isa, ix->n);
    dev_free(ix, &ix->d_kflag, ix->n);
  }
}

int ensure_staging(sapling_b200_index* ix) {
  if (ix->streams[0]) return 0;
  for (int i = 0; i < 3; i++) SB_CUDA_CHECK(cudaStreamCreateWithFlags(&ix->streams[i], cudaStreamNonBlocking));
  for (int i = 0; i < sapling_b200_index::kSlots; i++) {
    SB_CUDA_CHECK(cudaEventCreateWithFlags(&ix->ev_up[i], cudaEventDisableTiming));
    SB_CUDA_CHECK(cudaEventCreateWithFlags(&ix->ev_k[i], cudaEventDisableTiming));
    SB_CUDA_CHECK(cudaEventCreateWithFlags(&ix->ev_down[i], cudaEventDisableTiming));
    SB_CUDA_CHECK(cudaMalloc(&ix->d_in[i], sapling_b200_index::kChunk * 8));
    SB_CUDA_CHECK(cudaMalloc(&ix->d_raw[i], sapling_b200_index::kChunk * 8));
    SB_CUDA_CHECK(cudaMalloc(&ix->d_out[i], sapling_b200_index::kChunk * 8));
  }
  return 0;
}

int ensure_pinned(sapling_b200_index* ix) {
  if (ix->h_in[0]) return 0;
  for (int i = 0; i < sapling_b200_index::kSlots; i++) {
    SB_CUDA_CHECK(cudaMallocHost(&ix->h_in[i], sapling_b200_index::kChunk * 8));
    SB_CUDA_CHECK(cudaMallocHost(&ix->h_out[i], sapling_b200_index::kChunk * 8));
  }
  return 0;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// 2-bit pack of a query string the way the kernels read it (32 bases per word, base 0 in the top two bits).  Returns
// false when the string holds a byte other than A/C/G/T: the reference compares raw bytes (getLcp :118, :143), a 2-bit
// code cannot reproduce that ordering, so such queries are rejected rather than answered differently.
bool pack_query_string(const char* q, size_t slen, uint64_t* w) {
  bool ok = true;
  for (size_t j = 0; j < slen; j++) {
    const char c = q[j];
    uint64_t v = 0;
    if (c == 'C') v = 1;
    else if (c == 'G') v = 2;
    else if (c == 'T') v = 3;
    else if (c != 'A') ok = false;
    w[j >> 5] |= v << (62 - 2 * (j & 31));
  }
  return ok;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* sapling_b200_last_error(void) { return last_error(); }
const char* sapling_b200_version(void) { return "sapling_b200 0.2 (sm_100a)"; }

int64_t sapling_b200_kmerize(int k, const char* s) {
  // sapling_api.h:73-78 with vals[] of :494-498 (non-ACGT bytes hash as 0)
  uint64_t h = 0;
  for (int i = 0; i < k; i++) {
    const char c = s[i];
    const uint64_t v = c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0;
    h = (h << 2) | v;
  }
  return (int64_t)h;
}

int64_t sapling_b200_kmerize_adjusted(int k, int length, const char* s) {
  // sapling_api.h:83-90
  if (length >= k) return sapling_b200_kmerize(k, s);
  uint64_t h = (uint64_t)sapling_b200_kmerize(length, s);
  h = (h << 2) | 2u;
  return (int64_t)(h << (2 * (k - length - 1)));
}

static sapling_b200_index* new_index(unsigned flags, int nb, int maxMem, int k) {
  int dev = 0;
  if (require_device(&dev)) return nullptr;
  sapling_b200_index* ix = new sapling_b200_index();
  ix->device = dev;
  ix->flags = flags;
  ix->tune = Tuning::from_env();
  ix->nb = nb;                          // sapling_api.h:500
  if (k != -1) ix->k = k;               // :503-506
  if (maxMem != -1) ix->maxMem = maxMem;  // :507-510
  return ix;
}

sapling_b200_index* sapling_b200_open_cache(const char* path, unsigned flags);
int sapling_b200_save_cache(const sapling_b200_index* ix, const char* path);

sapling_b200_index* sapling_b200_open(const char* ref_fn, const char* sa_fn, const char* sap_fn, int nb, int maxMem,
                                      int k, const char* err_fn, unsigned flags) {
  // SAPLING_B200_CACHE=1: keep a private cache next to the .sap file (<sapFn>.b200) and open from it when it is there and
  // was built with the same k (and nb, when one is asked for).  Opt-in: the cache is not checked against the FASTA.
  const char* ce = getenv("SAPLING_B200_CACHE");
  const bool use_cache = ce && atoi(ce) != 0 && sap_fn && sap_fn[0] && !(err_fn && err_fn[0]);
  const std::string cache_fn = use_cache ? std::string(sap_fn) + ".b200" : std::string();
  if (use_cache && file_exists(cache_fn.c_str())) {
    sapling_b200_index* c = sapling_b200_open_cache(cache_fn.c_str(), flags);
    if (c && c->k == (k == -1 ? 21 : k) && (nb == -1 || c->nb == nb)) return c;
    if (c) delete c;  // another k / nb: rebuild from the reference's files below and replace the cache
  }
  sapling_b200_index* ix = new_index(flags, nb, maxMem, k);
  if (!ix) return nullptr;
  Say say(flags);
  auto fail = [&]() { delete ix; return (sapling_b200_index*)nullptr; };

  say("Reading reference genome\n");
  FILE* f = fopen(ref_fn, "rb");
  if (!f) { set_error("cannot open genome file %s", ref_fn); return fail(); }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::string text((size_t)sz, '\0');
  const size_t got = fread(&text[0], 1, (size_t)sz, f);
  fclose(f);
  clean_fasta(text.data(), got, &ix->genome, &ix->chr_ends);
  text.clear();
  text.shrink_to_fit();
  ix->n = ix->genome.size();
  if (validate_params(ix)) return fail();
  if (upload_genome(ix, ix->genome.data(), ix->n)) return fail();

  const bool have_sa = file_exists(sa_fn);
  const bool have_sap = file_exists(sap_fn);
  if (have_sa) {
    say("Reading suffix array from file\n");
    if (read_sa_file(ix, sa_fn)) return fail();
    say("Loaded suffix array of size %llu\n", (unsigned long long)ix->n);
  }
  bool model_given = false;
  if (have_sap) {
    say("Reading Sapling from file\n");
    std::vector<int64_t> xs, ys;
    if (read_sap_file(ix, sap_fn, &xs, &ys)) return fail();
    if (upload_model(ix, xs.data(), ys.data())) return fail();
    model_given = true;
  }
  if (build_missing(ix, model_given, err_fn, say)) return fail();
  if (!have_sa && sa_fn && sa_fn[0]) {
    say("Writing suffix array to file\n");
    if (write_sa_file(ix, sa_fn)) return fail();
  }
  if (!model_given && sap_fn && sap_fn[0]) {
    say("Writing Sapling to file\n");
    if (sapling_b200_write_sap(ix, sap_fn)) return fail();
  }
  if (use_cache && sapling_b200_save_cache(ix, cache_fn.c_str())) say("%s\n", last_error());  // not fatal
  drop_build_arrays(ix);
  return ix;
}

// ---- private index cache ---------------------------------------------------------------------------------
// The reference's files are what they are -- a .sa file is 16 bytes per base (50 GB at 3.1 Gbp) and has to be inverted,
// the FASTA has to be cleaned and packed.  The cache holds the arrays the device needs in the form it needs them
// (2-bit genome, 32-bit suffix array, {x, y} model: 5.6 bytes per base at c3 instead of 17 + FASTA) plus the scalars, so
// a later open is three sequential reads through pinned double buffers.  Native endianness, like the reference's files.
namespace {
constexpr char kCacheMagic[8] = {'S', 'B', '2', '0', '0', 'I', 'D', 'X'};
constexpr uint32_t kCacheVersion = 1;
struct CacheHeader {
  char magic[8];
  uint32_t version, k, nb, maxMem;
  int32_t five[5];
  uint32_t n_chr;
  uint64_t n, perfect, n_over, n_under, genome_words, model_entries;
};
constexpr size_t kIoPiece = (size_t)64 << 20;
int write_dev(FILE* f, const void* d, size_t bytes) {
  PinnedPair pp;
  if (pp.init(std::min(bytes ? bytes : 1, kIoPiece))) return -1;
  // download of piece i+1 overlaps the fwrite of piece i
  size_t o = 0;
  int slot = 0;
  size_t m = std::min(pp.bytes, bytes);
  if (m) {
    SB_CUDA_CHECK(cudaMemcpyAsync(pp.h[0], d, m, cudaMemcpyDeviceToHost, 0));
    cudaEventRecord(pp.ev[0], 0);
  }
  while (o < bytes) {
    const size_t next_o = o + m;
    const size_t next_m = next_o < bytes ? std::min(pp.bytes, bytes - next_o) : 0;
    if (next_m) {
      SB_CUDA_CHECK(cudaMemcpyAsync(pp.h[slot ^ 1], static_cast<const char*>(d) + next_o, next_m, cudaMemcpyDeviceToHost, 0));
      cudaEventRecord(pp.ev[slot ^ 1], 0);
    }
    SB_CUDA_CHECK(cudaEventSynchronize(pp.ev[slot]));
    if (fwrite(pp.h[slot], 1, m, f) != m) { set_error("index cache: short write"); return -1; }
    o = next_o;
    m = next_m;
    slot ^= 1;
  }
  return 0;
}
int read_dev(FILE* f, void* d, size_t bytes) {
  PinnedPair pp;
  if (pp.init(std::min(bytes ? bytes : 1, kIoPiece))) return -1;
  int slot = 0;
  for (size_t o = 0; o < bytes; o += pp.bytes, slot ^= 1) {
    const size_t m = std::min(pp.bytes, bytes - o);
    cudaEventSynchronize(pp.ev[slot]);  // the upload that last used this buffer
    if (fread(pp.h[slot], 1, m, f) != m) { set_error("index cache: file is truncated"); return -1; }
    SB_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(d) + o, pp.h[slot], m, cudaMemcpyHostToDevice, 0));
    cudaEventRecord(pp.ev[slot], 0);
  }
  SB_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}
// xlist / ylist as the wide table, from whichever form is resident, into a scratch device array
int wide_model_scratch(const sapling_b200_index* ix, ModelEntry** out) {
  const uint64_t count = (1ull << ix->nb) + 1;
  SB_CUDA_CHECK(cudaMalloc(out, count * sizeof(ModelEntry)));
  if (ix->d_model) {
    SB_CUDA_CHECK(cudaMemcpy(*out, ix->d_model, count * sizeof(ModelEntry), cudaMemcpyDeviceToDevice));
    return 0;
  }
  // widen into two int64 arrays, then interleave on the host side of nothing: do it in pieces through x / y scratch
  long long *dx = nullptr, *dy = nullptr;
  const uint64_t P = 1ull << 24;
  SB_CUDA_CHECK(cudaMalloc(&dx, std::min(P, count) * 8));
  if (cudaMalloc(&dy, std::min(P, count) * 8) != cudaSuccess) { cudaFree(dx); set_error("model scratch allocation failed"); return -1; }
  int rc = 0;
  for (uint64_t o = 0; o < count && !rc; o += P) {
    const uint64_t c = std::min(P, count - o);
    rc = widen_model(ix->d_narrow, ix->nb, 2 * ix->k - ix->nb, ix->last_x, ix->last_y, o, c, dx, dy, 0);
    if (!rc && cudaMemcpy2D(reinterpret_cast<char*>(*out + o), 16, dx, 8, 8, c, cudaMemcpyDeviceToDevice) != cudaSuccess) rc = -1;
    if (!rc && cudaMemcpy2D(reinterpret_cast<char*>(*out + o) + 8, 16, dy, 8, 8, c, cudaMemcpyDeviceToDevice) != cudaSuccess) rc = -1;
  }
  cudaFree(dx);
  cudaFree(dy);
  if (rc) { cudaFree(*out); *out = nullptr; set_error("model reconstruction failed"); }
  return rc;
}
}  // namespace

int sapling_b200_save_cache(const sapling_b200_index* ix, const char* path) {
  if (!ix || !path || !path[0]) { set_error("save_cache: null index or empty path"); return -1; }
  cudaSetDevice(ix->device);
  SaGuard sg;
  if (sg.get(ix)) return -1;
  ModelEntry* d_wide = nullptr;
  if (wide_model_scratch(ix, &d_wide)) return -1;
  FILE* f = fopen(path, "wb");
  if (!f) { cudaFree(d_wide); set_error("cannot write index cache %s", path); return -1; }
  CacheHeader h{};
  memcpy(h.magic, kCacheMagic, 8);
  h.version = kCacheVersion;
  h.k = (uint32_t)ix->k; h.nb = (uint32_t)ix->nb; h.maxMem = (uint32_t)ix->maxMem;
  h.five[0] = ix->stats.maxOver; h.five[1] = ix->stats.maxUnder; h.five[2] = ix->stats.meanError;
  h.five[3] = ix->stats.mostOver; h.five[4] = ix->stats.mostUnder;
  h.n_chr = (uint32_t)ix->chr_ends.size();
  h.n = ix->n; h.perfect = ix->stats.perfect; h.n_over = ix->stats.nOver; h.n_under = ix->stats.nUnder;
  h.genome_words = packed_words(ix->n);
  h.model_entries = (1ull << ix->nb) + 1;
  int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? 0 : -1;
  for (const auto& ce : ix->chr_ends) {
    const uint32_t len = (uint32_t)ce.second.size();
    if (fwrite(&ce.first, 8, 1, f) != 1 || fwrite(&len, 4, 1, f) != 1 || (len && fwrite(ce.second.data(), 1, len, f) != len)) rc = -1;
  }
  if (rc) set_error("index cache: short write");
  if (!rc) rc = write_dev(f, ix->d_genome, h.genome_words * 8);
  if (!rc) rc = write_dev(f, sg.sa, ix->n * 4);
  if (!rc) rc = write_dev(f, d_wide, h.model_entries * sizeof(ModelEntry));
  if (!rc && fwrite(kCacheMagic, 8, 1, f) != 1) { set_error("index cache: short write"); rc = -1; }
  if (fclose(f) != 0 && !rc) { set_error("index cache: close failed"); rc = -1; }
  cudaFree(d_wide);
  if (rc) remove(path);
  return rc;
}

sapling_b200_index* sapling_b200_open_cache(const char* path, unsigned flags) {
  FILE* f = path ? fopen(path, "rb") : nullptr;
  if (!f) { set_error("cannot open index cache %s", path ? path : "(null)"); return nullptr; }
  CacheHeader h{};
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kCacheMagic, 8) != 0 || h.version != kCacheVersion) {
    set_error("%s is not a sapling_b200 index cache (or another version of it)", path);
    fclose(f);
    return nullptr;
  }
  sapling_b200_index* ix = new_index(flags, (int)h.nb, (int)h.maxMem, (int)h.k);
  if (!ix) { fclose(f); return nullptr; }
  Say say(flags);
  auto fail = [&]() { fclose(f); delete ix; return (sapling_b200_index*)nullptr; };
  ix->n = h.n;
  if (validate_params(ix)) return fail();
  if (h.genome_words != packed_words(h.n) || h.model_entries != (1ull << h.nb) + 1 || h.nb < 1 || h.nb > 31) {
    set_error("index cache %s: inconsistent header", path);
    return fail();
  }
  ix->stats.maxOver = h.five[0]; ix->stats.maxUnder = h.five[1]; ix->stats.meanError = h.five[2];
  ix->stats.mostOver = h.five[3]; ix->stats.mostUnder = h.five[4];
  ix->stats.perfect = h.perfect; ix->stats.nOver = h.n_over; ix->stats.nUnder = h.n_under;
  for (uint32_t i = 0; i < h.n_chr; i++) {
    uint64_t end = 0;
    uint32_t len = 0;
    if (fread(&end, 8, 1, f) != 1 || fread(&len, 4, 1, f) != 1 || len > (1u << 20)) { set_error("index cache: bad chromosome table"); return fail(); }
    std::string name(len, '\0');
    if (len && fread(&name[0], 1, len, f) != len) { set_error("index cache: file is truncated"); return fail(); }
    ix->chr_ends.emplace_back(end, name);
  }
  say("Reading index cache\n");
  if (dev_alloc(ix, &ix->d_genome, h.genome_words) || read_dev(f, ix->d_genome, h.genome_words * 8)) return fail();
  if (dev_alloc(ix, &ix->d_sa, h.n) || read_dev(f, ix->d_sa, h.n * 4)) return fail();
  if (dev_alloc(ix, &ix->d_model, h.model_entries) || read_dev(f, ix->d_model, h.model_entries * sizeof(ModelEntry))) return fail();
  char tail[8];
  if (fread(tail, 8, 1, f) != 1 || memcmp(tail, kCacheMagic, 8) != 0) { set_error("index cache %s: file is truncated", path); return fail(); }
  {  // Sapling::reference for the callers that read it (sapling_b200_genome)
    char* d_ascii = nullptr;
    if (cudaMalloc(&d_ascii, h.n) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc ascii genome failed"); return fail(); }
    ix->genome.resize(h.n);
    int rc = unpack_genome(ix->d_genome, h.n, d_ascii, 0);
    if (!rc && cudaMemcpy(&ix->genome[0], d_ascii, h.n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
    cudaFree(d_ascii);
    if (rc) { set_error("index cache: genome download failed"); return fail(); }
  }
  if (build_missing(ix, true, nullptr, say)) return fail();
  drop_build_arrays(ix);
  fclose(f);
  return ix;
}

static sapling_b200_index* create_common(const char* genome, uint64_t n, const uint32_t* sa, int nb, int maxMem,
                                         int k, const int64_t* xlist, const int64_t* ylist, const int* five,
                                         unsigned flags) {
  sapling_b200_index* ix = new_index(flags, nb, maxMem, k);
  if (!ix) return nullptr;
  Say say(flags | SAPLING_B200_QUIET);
  auto fail = [&]() { delete ix; return (sapling_b200_index*)nullptr; };
  ix->n = n;
  ix->genome.assign(genome, n);
  if (validate_params(ix)) return fail();
  if (upload_genome(ix, genome, n)) return fail();
  if (sa) {
    if (dev_alloc(ix, &ix->d_sa, n)) return fail();
    if (cudaMemcpy(ix->d_sa, sa, n * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("suffix array upload failed");
      return fail();
    }
  }
  bool model_given = false;
  if (xlist && ylist && five) {
    ix->stats.maxOver = five[0]; ix->stats.maxUnder = five[1]; ix->stats.meanError = five[2];
    ix->stats.mostOver = five[3]; ix->stats.mostUnder = five[4];
    if (ix->nb < 1 || ix->nb > 31) { set_error("nb must be given with a model"); return fail(); }
    if (upload_model(ix, xlist, ylist)) return fail();
    model_given = true;
  }
  if (build_missing(ix, model_given, nullptr, say)) return fail();
  drop_build_arrays(ix);
  return ix;
}

sapling_b200_index* sapling_b200_create(const char* genome, uint64_t n, const uint32_t* sa, int nb, int maxMem, int k,
                                        unsigned flags) {
  return create_common(genome, n, sa, nb, maxMem, k, nullptr, nullptr, nullptr, flags);
}

sapling_b200_index* sapling_b200_create_with_model(const char* genome, uint64_t n, const uint32_t* sa, int k, int nb,
                                                   const int64_t* xlist, const int64_t* ylist, const int* five,
                                                   unsigned flags) {
  return create_common(genome, n, sa, nb, -1, k, xlist, ylist, five, flags);
}

sapling_b200_index* sapling_b200_create_synthetic(uint64_t seed, uint64_t n, int nb, int maxMem, int k,
                                                  int keep_host_genome, unsigned flags) {
  sapling_b200_index* ix = new_index(flags, nb, maxMem, k);
  if (!ix) return nullptr;
  Say say(flags);
  auto fail = [&]() { delete ix; return (sapling_b200_index*)nullptr; };
  ix->n = n;
  if (validate_params(ix)) return fail();
  if (dev_alloc(ix, &ix->d_genome, packed_words(n))) return fail();
  if (synth_genome_packed(seed, n, ix->d_genome, 0)) return fail();
  if (keep_host_genome) {
    char* d_ascii = nullptr;
    if (cudaMalloc(&d_ascii, n) != cudaSuccess) { set_error("cudaMalloc ascii genome failed"); return fail(); }
    ix->genome.resize(n);
    int rc = unpack_genome(ix->d_genome, n, d_ascii, 0);
    cudaError_t e = cudaMemcpy(&ix->genome[0], d_ascii, n, cudaMemcpyDeviceToHost);
    cudaFree(d_ascii);
    if (rc || e != cudaSuccess) { set_error("genome download failed"); return fail(); }
  }
  ix->chr_ends.push_back({n, "chr1"});
  if (build_missing(ix, false, nullptr, say)) return fail();
  drop_build_arrays(ix);
  return ix;
}

// ---- replicas on other GPUs ---------------------------------------------------------------------------------
// The index is ingested / built ONCE, on the handle's own GPU; every other GPU named in gpu_mask then receives a copy of the
// resident arrays by cudaMemcpyPeer (NVLink / NVSwitch when the GPUs are peers, through the host otherwise).  There is no
// collective on the query path: the host-pointer batch calls cut a batch into contiguous slices, one per GPU.
int sapling_b200_replicate(sapling_b200_index* ix, uint64_t gpu_mask) {
  if (!ix) { set_error("null index"); return -1; }
  if (ix->is_replica) { set_error("replicate: handle is itself a replica"); return -1; }
  int ndev = 0;
  SB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  std::lock_guard<std::mutex> lock(ix->mu);
  const uint64_t B = 1ull << ix->nb;
  std::vector<sapling_b200_index*> fresh;
  auto abandon = [&](int dev, const char* what) {
    set_error("replicate to device %d: %s: %s", dev, what, cudaGetErrorString(cudaGetLastError()));
    for (auto* r : fresh) {
      cudaSetDevice(r->device);
      cudaDeviceSynchronize();
      delete r;
    }
    cudaSetDevice(ix->device);
    return -1;
  };
  // pass 1: allocate on every new GPU and start its copies (asynchronous, one stream per destination: the copies to
  // different GPUs run side by side through the switch)
  for (int dev = 0; dev < ndev && dev < 64; dev++) {
    if (!((gpu_mask >> dev) & 1ull) || dev == ix->device) continue;
    bool have = false;
    for (auto* r : ix->replicas) have |= r->device == dev;
    if (have) continue;
    if (check_device(dev)) return abandon(dev, "device check");
    sapling_b200_index* r = new sapling_b200_index();
    fresh.push_back(r);
    r->device = dev;
    r->is_replica = true;
    r->flags = ix->flags;
    r->tune = ix->tune;
    r->n = ix->n; r->k = ix->k; r->nb = ix->nb; r->maxMem = ix->maxMem;
    r->stats = ix->stats;
    r->line_bases = ix->line_bases;
    r->last_x = ix->last_x; r->last_y = ix->last_y;
    r->hints = ix->hints;
    // direct NVLink / NVSwitch copies need peer access enabled on both ends; "already enabled" is fine
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, ix->device, dev) == cudaSuccess && can) {
      cudaSetDevice(ix->device);
      cudaDeviceEnablePeerAccess(dev, 0);
      cudaGetLastError();
    }
    if (cudaSetDevice(dev) != cudaSuccess) return abandon(dev, "cudaSetDevice");
    if (cudaDeviceCanAccessPeer(&can, dev, ix->device) == cudaSuccess && can) {
      cudaDeviceEnablePeerAccess(ix->device, 0);
      cudaGetLastError();
    }
    auto copy = [&](auto** dst, const auto* src, uint64_t count) -> bool {
      if (!src) return true;
      if (dev_alloc(r, dst, count)) return false;
      return cudaMemcpyPeerAsync(*dst, dev, src, ix->device, count * sizeof(**dst), 0) == cudaSuccess;
    };
    if (!copy(&r->d_genome, ix->d_genome, packed_words(ix->n))) return abandon(dev, "genome");
    if (!copy(&r->d_lines, ix->d_lines, line_sectors(ix->n) * 8)) return abandon(dev, "rank lines");
    if (!copy(&r->d_narrow, ix->d_narrow, B + 1)) return abandon(dev, "model");
    if (!copy(&r->d_model, ix->d_model, B + 1)) return abandon(dev, "wide model");
    if (!copy(&r->d_isa, ix->d_isa, ix->n)) return abandon(dev, "inverse suffix array");
    if (!copy(&r->d_kflag, ix->d_kflag, ix->n)) return abandon(dev, "k flags");
    if (dev_alloc(r, &r->d_oob, 1) || cudaMemsetAsync(r->d_oob, 0, 8, 0) != cudaSuccess) return abandon(dev, "counter");
  }
  // pass 2: wait for them
  for (auto* r : fresh) {
    if (cudaSetDevice(r->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return abandon(r->device, "copy");
  }
  for (auto* r : fresh) ix->replicas.push_back(r);
  cudaSetDevice(ix->device);
  return 0;
}

sapling_b200_index* sapling_b200_open_multi(const char* ref_fn, const char* sa_fn, const char* sap_fn, int nb, int maxMem,
                                            int k, const char* err_fn, unsigned flags, uint64_t gpu_mask) {
  sapling_b200_index* ix = sapling_b200_open(ref_fn, sa_fn, sap_fn, nb, maxMem, k, err_fn, flags);
  if (ix && sapling_b200_replicate(ix, gpu_mask)) { delete ix; return nullptr; }
  return ix;
}

int sapling_b200_num_devices(const sapling_b200_index* ix) { return ix ? 1 + (int)ix->replicas.size() : 0; }

void sapling_b200_close(sapling_b200_index* ix) { delete ix; }

int sapling_b200_info(const sapling_b200_index* ix, uint64_t* n, int* k, int* nb, int* maxOver, int* maxUnder,
                      int* meanError, int* mostOver, int* mostUnder) {
  if (!ix) { set_error("null index"); return -1; }
  if (n) *n = ix->n;
  if (k) *k = ix->k;
  if (nb) *nb = ix->nb;
  if (maxOver) *maxOver = ix->stats.maxOver;
  if (maxUnder) *maxUnder = ix->stats.maxUnder;
  if (meanError) *meanError = ix->stats.meanError;
  if (mostOver) *mostOver = ix->stats.mostOver;
  if (mostUnder) *mostUnder = ix->stats.mostUnder;
  return 0;
}

const char* sapling_b200_genome(const sapling_b200_index* ix) {
  return (ix && !ix->genome.empty()) ? ix->genome.c_str() : nullptr;
}
size_t sapling_b200_num_chr(const sapling_b200_index* ix) { return ix ? ix->chr_ends.size() : 0; }
uint64_t sapling_b200_chr(const sapling_b200_index* ix, size_t i, const char** name) {
  if (!ix || i >= ix->chr_ends.size()) return 0;
  if (name) *name = ix->chr_ends[i].second.c_str();
  return ix->chr_ends[i].first;
}

int sapling_b200_build_stats(const sapling_b200_index* ix, uint64_t* perfect, uint64_t* n_over, uint64_t* n_under) {
  if (!ix) { set_error("null index"); return -1; }
  if (perfect) *perfect = ix->stats.perfect;
  if (n_over) *n_over = ix->stats.nOver;
  if (n_under) *n_under = ix->stats.nUnder;
  return 0;
}

int sapling_b200_model(const sapling_b200_index* ix, int64_t* xlist, int64_t* ylist) {
  if (!ix) { set_error("null index"); return -1; }
  const uint64_t count = (1ull << ix->nb) + 1;
  cudaSetDevice(ix->device);
  if (ix->d_model) {
    std::vector<ModelEntry> m(count);
    SB_CUDA_CHECK(cudaMemcpy(m.data(), ix->d_model, count * sizeof(ModelEntry), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < count; i++) {
      if (xlist) xlist[i] = m[i].x;
      if (ylist) ylist[i] = m[i].y;
    }
    return 0;
  }
  // rebuilt from the narrow table, piece by piece
  const uint64_t P = std::min<uint64_t>(count, 1ull << 24);
  long long *dx = nullptr, *dy = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&dx, P * 8));
  if (cudaMalloc(&dy, P * 8) != cudaSuccess) { cudaFree(dx); set_error("model: scratch allocation failed"); return -1; }
  int rc = 0;
  for (uint64_t o = 0; o < count && !rc; o += P) {
    const uint64_t c = std::min(P, count - o);
    rc = widen_model(ix->d_narrow, ix->nb, 2 * ix->k - ix->nb, ix->last_x, ix->last_y, o, c, dx, dy, 0);
    if (!rc && xlist && cudaMemcpy(xlist + o, dx, c * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
    if (!rc && ylist && cudaMemcpy(ylist + o, dy, c * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
  }
  cudaFree(dx);
  cudaFree(dy);
  if (rc) set_error("model: reconstruction from the narrow table failed");
  return rc;
}

int sapling_b200_rev(const sapling_b200_index* ix, uint64_t first, uint64_t count, uint32_t* out) {
  if (!ix || first + count > ix->n) { set_error("rev: range out of bounds"); return -1; }
  if (count == 0) return 0;
  cudaSetDevice(ix->device);
  if (ix->d_sa) {
    SB_CUDA_CHECK(cudaMemcpy(out, ix->d_sa + first, count * 4, cudaMemcpyDeviceToHost));
    return 0;
  }
  const uint64_t P = std::min<uint64_t>(count, 1ull << 26);
  uint32_t* d = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d, P * 4));
  int rc = 0;
  for (uint64_t o = 0; o < count && !rc; o += P) {
    const uint64_t c = std::min(P, count - o);
    rc = launch_rev_extract(ix->view(), first + o, c, d, 0);
    if (!rc && cudaMemcpy(out + o, d, c * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("rev: download failed"); rc = -1; }
  }
  cudaFree(d);
  return rc;
}

int sapling_b200_sa_rank(const sapling_b200_index* ix, uint64_t first, uint64_t count, uint32_t* out) {
  if (!ix || first + count > ix->n) { set_error("sa_rank: range out of bounds"); return -1; }
  if (!ix->d_isa) { set_error("sa_rank: ISA not resident (open with SAPLING_B200_KEEP_BUILD)"); return -1; }
  cudaSetDevice(ix->device);
  SB_CUDA_CHECK(cudaMemcpy(out, ix->d_isa + first, count * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int sapling_b200_write_sap(const sapling_b200_index* ix, const char* path) {
  if (!ix) { set_error("null index"); return -1; }
  const uint64_t count = (1ull << ix->nb) + 1;
  std::vector<int64_t> xs(count), ys(count);
  if (sapling_b200_model(ix, xs.data(), ys.data())) return -1;
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("cannot write %s", path); return -1; }
  // sapling_api.h:656-674
  fwrite(&ix->nb, sizeof(int), 1, f);
  if (ix->nb <= 30) { const int c32 = (int)count; fwrite(&c32, sizeof(int), 1, f); }
  else fwrite(&count, 8, 1, f);
  fwrite(xs.data(), 8, count, f);
  fwrite(ys.data(), 8, count, f);
  const int five[5] = {ix->stats.maxOver, ix->stats.maxUnder, ix->stats.meanError, ix->stats.mostOver,
                       ix->stats.mostUnder};
  fwrite(five, sizeof(int), 5, f);
  fclose(f);
  return 0;
}

int sapling_b200_write_sa(const sapling_b200_index* ix, const char* path) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  return write_sa_file(ix, path);
}

int sapling_b200_check_sa(const sapling_b200_index* ix, uint32_t max_chars, uint64_t* bad_order, uint64_t* undecided,
                          uint64_t* bad_perm) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  SaGuard sg;
  if (sg.get(ix)) return -1;
  return check_suffix_array(ix->d_genome, ix->n, sg.sa, ix->d_isa, max_chars, bad_order, undecided, bad_perm, 0);
}

uint64_t sapling_b200_device_bytes(const sapling_b200_index* ix) { return ix ? ix->device_bytes : 0; }

uint64_t sapling_b200_launch_count(const sapling_b200_index* ix) {
  if (!ix) return 0;
  uint64_t v = ix->launches.load(std::memory_order_relaxed);
  for (auto* r : ix->replicas) v += r->launches.load(std::memory_order_relaxed);
  return v;
}

// ---------------------------------------------------------------------------------------------

// How many top k-mer bits to partition a batch of nq queries by; 0 = answer it in the caller's order.
// The slice of the index one bin maps to (rank lines + model, both monotone in the k-mer) should fit the part of L2 that
// data shared by all SMs gets (~48 MB, profiles/r1_experiments.md section 1) with room for the genome and the streams;
// each bin should still receive enough queries to amortise its line fills.
static int partition_bits(const sapling_b200_index* ix, size_t nq) {
  const Tuning& t = ix->tune;
  if (!t.partition) return 0;
  if (nq < t.partition_min || nq >= (1ull << 32)) return 0;
  const int kbits = 2 * ix->k;
  int bits;
  if (t.partition_bits >= 0) {
    bits = t.partition_bits;
  } else {
    const double sa_bytes = (double)line_sectors(ix->n) * 32.0;
    const double model_bytes = (ix->d_narrow ? 8.0 : 16.0) * (double)(1ull << ix->nb);
    // 32 MB slices, but no more than 512 of them: a (chunk, slice) run of the scatter and un-permute passes then still
    // averages 32 queries = 256 bytes.  Measured at c3 (gpurun s13 / s14, ms per 250 M queries, passes + kernel Q): 8 bits
    // (105 MB slices) 2.30 + 5.92, 9 bits (53 MB) 2.59 + 5.66, 10 bits (26 MB) 3.14 + 5.60, 11 bits 4.9 + 5.6; at c2 4 to 8
    // bits differ by 3 %.
    const double slice = 32e6;
    const int auto_max = 9;
    bits = 1;
    while (bits < auto_max && (sa_bytes + model_bytes) / (double)(1ull << bits) > slice) bits++;
    while (bits > 0 && (nq >> bits) < 4096) bits--;
    if (bits < 3) return 0;
  }
  if (bits > kPartMaxBits) bits = kPartMaxBits;
  if (bits > kbits) bits = kbits;
  return bits < 1 ? 0 : bits;
}

// One batch of k-mers already on the device -> answers (long long, or uint32_t when d_out32 != nullptr), enqueued on st.
// An unpartitioned batch takes the in-order kernel (query.cu) when every warp of its grid gets a few tiles to pipeline.
static bool inorder_batch(const sapling_b200_index* ix, size_t nq) {
  return ix->tune.inorder_min >= 0 && nq >= (size_t)ix->tune.inorder_min && nq < (1ull << 32);
}

static int run_kmer_batch(sapling_b200_index* ix, const IndexView& v, const uint64_t* d_kmers, size_t nq, long long* d_out,
                          uint32_t* d_out32, cudaStream_t st) {
  if (nq == 0) return 0;
  const int bits = partition_bits(ix, nq);
  cudaEvent_t evs[5];
  cudaEvent_t* ev = nullptr;
  {
    std::lock_guard<std::mutex> lock(ix->mu_prof);
    if (ix->profiling) {
      sapling_b200_index::ProfCall c;
      for (int i = 0; i < 5; i++) {
        if (cudaEventCreate(&c.ev[i]) != cudaSuccess) {
          for (int j = 0; j < i; j++) cudaEventDestroy(c.ev[j]);
          set_error("cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
          return -1;
        }
        evs[i] = c.ev[i];  // copied while the list is locked
      }
      c.partitioned = bits != 0;
      ix->prof_calls.push_back(c);
      ev = evs;
    }
  }
  // per-stream scratch: the partition workspace, or just the tile counter of the in-order kernel
  auto scratch = [&](size_t need) -> void* {
    std::lock_guard<std::mutex> lock(ix->mu_ws);
    sapling_b200_index::PartWs& w = ix->part_ws[st];
    if (w.bytes < need) {
      if (w.p) { cudaFree(w.p); ix->device_bytes -= w.bytes; }  // cudaFree waits for work that still uses the block
      w.p = nullptr;
      w.bytes = 0;
      if (cudaMalloc(&w.p, need) != cudaSuccess) {
        cudaGetLastError();
        w.p = nullptr;  // no room for the scratch: the plain kernel needs none
      } else {
        w.bytes = need;
        ix->device_bytes += need;
      }
    }
    return w.p;
  };
  // a batch answered in the caller's order: the pipelined in-order kernel from a few tiles per warp on, else one query
  // per thread
  auto plain = [&]() -> int {
    ix->launches.fetch_add(1, std::memory_order_relaxed);
    if (ev) {  // unpartitioned call: stages 0, 1 and 3 are empty
      cudaEventRecord(ev[0], st);
      cudaEventRecord(ev[1], st);
      cudaEventRecord(ev[2], st);
    }
    void* tiles = inorder_batch(ix, nq) ? scratch(256) : nullptr;
    const int rc = tiles ? launch_kmer_query_inorder(v, d_kmers, nq, d_out, d_out32, static_cast<unsigned long long*>(tiles), st)
                         : launch_kmer_query(v, d_kmers, nq, d_out, d_out32, ix->tune.occupancy, st);
    if (ev) {
      cudaEventRecord(ev[3], st);
      cudaEventRecord(ev[4], st);
    }
    return rc;
  };
  if (bits == 0) return plain();
  void* ws = scratch(partition_workspace_bytes(nq, bits));
  if (!ws) return plain();
  ix->launches.fetch_add(8, std::memory_order_relaxed);  // histogram, three column-scan passes, bin scan, scatter, query, un-permute
  return launch_partitioned_query(v, d_kmers, nq, d_out, d_out32, ws, bits, st, ev);
}

int sapling_b200_profile(sapling_b200_index* ix, int on) {
  if (!ix) { set_error("null index"); return -1; }
  std::lock_guard<std::mutex> lock(ix->mu_prof);
  ix->profiling = on != 0;
  return 0;
}

int sapling_b200_stage_ms(sapling_b200_index* ix, double ms[4]) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  std::lock_guard<std::mutex> lock(ix->mu_prof);
  for (int i = 0; i < 4; i++) ms[i] = 0.0;
  int calls = 0;
  for (auto& c : ix->prof_calls) {
    if (cudaEventSynchronize(c.ev[4]) == cudaSuccess) {
      for (int i = 0; i < 4; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c.ev[i], c.ev[i + 1]) == cudaSuccess) ms[i] += (double)t;
      }
      calls++;
    }
    cudaGetLastError();
    for (int i = 0; i < 5; i++) cudaEventDestroy(c.ev[i]);
  }
  ix->prof_calls.clear();
  return calls;
}

int sapling_b200_query_partition_bits(const sapling_b200_index* ix, size_t nq) {
  return ix ? partition_bits(ix, nq) : 0;
}

const char* sapling_b200_query_kernel_for(const sapling_b200_index* ix, size_t nq, int* blocks_per_sm) {
  if (!ix) return "";
  const bool ordered = partition_bits(ix, nq) != 0 || inorder_batch(ix, nq);
  if (blocks_per_sm) *blocks_per_sm = kmer_query_blocks_per_sm(ordered, ix->tune.occupancy);
  return kmer_query_kernel_name(ordered);
}
const char* sapling_b200_query_kernel(const sapling_b200_index* ix, int* blocks_per_sm) {
  return sapling_b200_query_kernel_for(ix, 1, blocks_per_sm);
}

int sapling_b200_query_batch_dev(sapling_b200_index* ix, const uint64_t* d_kmers, size_t nq, int64_t* d_out,
                                 void* stream) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  return run_kmer_batch(ix, ix->view(), d_kmers, nq, reinterpret_cast<long long*>(d_out), nullptr,
                        static_cast<cudaStream_t>(stream));
}

int sapling_b200_query_batch_u32_dev(sapling_b200_index* ix, const uint64_t* d_kmers, size_t nq, uint32_t* d_out,
                                     void* stream) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  return run_kmer_batch(ix, ix->view(), d_kmers, nq, nullptr, d_out, static_cast<cudaStream_t>(stream));
}

// The chunk pipeline of the host-pointer entry points on ONE device.  kmers: a little-endian bit stream of nq integers of
// kmer_bits bits each (64: plain uint64; 40 / 48 / 56: whole bytes; 2k: nothing but the k-mer), starting on a byte boundary;
// out: nq answers of out_bytes (8: long long, 4: uint32_t with 0xFFFFFFFF for -1).
static int host_batch_one(sapling_b200_index* ix, const char* kmers, int kmer_bits, size_t nq, char* out, int out_bytes) {
  if (nq == 0) return 0;
  std::lock_guard<std::mutex> lock(ix->mu);
  if (cudaSetDevice(ix->device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", ix->device); return -1; }
  if (ensure_staging(ix)) return -1;
  const bool pin_in = is_pinned(kmers), pin_out = is_pinned(out);
  if ((!pin_in || !pin_out) && ensure_pinned(ix)) return -1;
  const IndexView v = ix->view();
  size_t CH = 1u << 18;  // ~sqrt(2 * nq * 1e5) rounded to a power of two, within [256 Ki, 4 Mi]
  while (CH < sapling_b200_index::kChunk && (double)(2 * CH) * (double)(2 * CH) <= 8.0 * (double)nq * 1e5) CH *= 2;
  if (ix->tune.chunk_log2 >= 12 && (1ull << ix->tune.chunk_log2) <= sapling_b200_index::kChunk) CH = 1ull << ix->tune.chunk_log2;
  const int NS = sapling_b200_index::kSlots;
  const size_t nchunks = (nq + CH - 1) / CH;
  cudaStream_t s_up = ix->streams[0], s_k = ix->streams[1], s_down = ix->streams[2];
  auto in_bytes = [&](size_t queries) { return (queries * (size_t)kmer_bits + 7) / 8; };  // CH is a multiple of 8 queries
  int rc = 0;
  // chunk c lives in slot c % NS: upload -> kernel -> download, each on its own stream, ordered by events
  for (size_t c = 0; c < nchunks + (size_t)NS && !rc; c++) {
    if (c >= (size_t)NS) {  // retire chunk c-NS: its slot is about to be reused
      const size_t r = c - (size_t)NS;
      const int s = (int)(r % (size_t)NS);
      if (cudaEventSynchronize(ix->ev_down[s]) != cudaSuccess) { rc = -1; break; }
      if (!pin_out) {
        const size_t o = r * CH, m = std::min(CH, nq - o);
        memcpy(out + o * out_bytes, ix->h_out[s], m * out_bytes);
      }
    }
    if (c < nchunks) {
      const int s = (int)(c % (size_t)NS);
      const size_t o = c * CH, m = std::min(CH, nq - o);
      const char* src = kmers + in_bytes(o);
      if (!pin_in) {
        memcpy(ix->h_in[s], src, in_bytes(m));
        src = static_cast<const char*>(ix->h_in[s]);
      }
      void* d_up = kmer_bits == 64 ? (void*)ix->d_in[s] : ix->d_raw[s];
      if (cudaMemcpyAsync(d_up, src, in_bytes(m), cudaMemcpyHostToDevice, s_up) != cudaSuccess) { rc = -1; break; }
      cudaEventRecord(ix->ev_up[s], s_up);
      cudaStreamWaitEvent(s_k, ix->ev_up[s], 0);
      if (kmer_bits != 64 && launch_unpack_kmers(ix->d_raw[s], kmer_bits, m, ix->d_in[s], s_k)) { rc = -1; break; }
      if (run_kmer_batch(ix, v, ix->d_in[s], m, out_bytes == 8 ? ix->d_out[s] : nullptr,
                         out_bytes == 8 ? nullptr : reinterpret_cast<uint32_t*>(ix->d_out[s]), s_k)) { rc = -2; break; }
      cudaEventRecord(ix->ev_k[s], s_k);
      cudaStreamWaitEvent(s_down, ix->ev_k[s], 0);
      void* dst = pin_out ? (void*)(out + o * out_bytes) : ix->h_out[s];
      if (cudaMemcpyAsync(dst, ix->d_out[s], m * out_bytes, cudaMemcpyDeviceToHost, s_down) != cudaSuccess) { rc = -1; break; }
      cudaEventRecord(ix->ev_down[s], s_down);
    }
  }
  if (rc) {
    // leave nothing in flight behind an error: copies may still be writing into the caller's buffer or the staging slots
    for (int i = 0; i < 3; i++) cudaStreamSynchronize(ix->streams[i]);
    if (rc == -1) set_error("query_batch: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return 0;  // every chunk was retired (event-synchronised) inside the loop
}

// Host-pointer batch over the handle's GPU and its replicas: contiguous slices, one host thread per device.  Slices start
// at multiples of 8 queries, which is a byte boundary of the bit stream for every kmer_bits.
static int host_batch(sapling_b200_index* ix, const void* kmers, int kmer_bits, size_t nq, void* out, int out_bytes) {
  if (!ix) { set_error("null index"); return -1; }
  if (nq == 0) return 0;
  const char* in = static_cast<const char*>(kmers);
  char* o = static_cast<char*>(out);
  const size_t G = 1 + ix->replicas.size();
  if (G == 1 || nq < 2 * G * 65536) return host_batch_one(ix, in, kmer_bits, nq, o, out_bytes);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++) {
    const size_t lo = (nq * g / G) & ~(size_t)7, hi = g + 1 == G ? nq : ((nq * (g + 1) / G) & ~(size_t)7);
    sapling_b200_index* dev_ix = g == 0 ? ix : ix->replicas[g - 1];
    th.emplace_back([=, &rcs, &errs]() {
      rcs[g] = host_batch_one(dev_ix, in + lo * (size_t)kmer_bits / 8, kmer_bits, hi - lo, o + lo * out_bytes, out_bytes);
      if (rcs[g]) errs[g] = last_error();  // the error text is thread-local
    });
  }
  for (auto& t : th) t.join();
  cudaSetDevice(ix->device);
  for (size_t g = 0; g < G; g++)
    if (rcs[g]) { set_error("%s", errs[g].c_str()); return -1; }
  return 0;
}

int sapling_b200_query_batch(sapling_b200_index* ix, const uint64_t* kmers, size_t nq, int64_t* out) {
  return host_batch(ix, kmers, 64, nq, out, 8);
}

int sapling_b200_query_batch_u32(sapling_b200_index* ix, const void* kmers, int kmer_bytes, size_t nq, uint32_t* out) {
  if (ix && (kmer_bytes < 1 || kmer_bytes > 8 || 8 * kmer_bytes < 2 * ix->k)) {
    set_error("query_batch_u32: kmer_bytes=%d cannot hold a %d-mer (need ceil(2k/8) .. 8)", kmer_bytes, ix->k);
    return -1;
  }
  return host_batch(ix, kmers, 8 * kmer_bytes, nq, out, 4);
}

int sapling_b200_query_batch_bits(sapling_b200_index* ix, const void* kmers, int kmer_bits, size_t nq, uint32_t* out) {
  if (ix && (kmer_bits < 2 * ix->k || kmer_bits > 64)) {
    set_error("query_batch_bits: kmer_bits=%d cannot hold a %d-mer (need 2k .. 64)", kmer_bits, ix->k);
    return -1;
  }
  return host_batch(ix, kmers, kmer_bits, nq, out, 4);
}

int sapling_b200_query_str_batch(sapling_b200_index* ix, const char* s, const uint64_t* offsets, const uint32_t* slens,
                                 const uint32_t* lengths, const int64_t* kmers, size_t nq, int64_t* out) {
  if (!ix) { set_error("null index"); return -1; }
  if (nq == 0) return 0;
  cudaSetDevice(ix->device);
  // pack every string to 2 bits/base, 32 bases per word, each query starting on a word boundary
  std::vector<uint64_t> word_off(nq);
  uint64_t total = 0;
  for (size_t i = 0; i < nq; i++) {
    if (lengths && lengths[i] > slens[i]) { set_error("query %zu: length > s.length() reads past the string in the reference", i); return -1; }
    word_off[i] = total;
    total += (slens[i] + 31) / 32 + 1;
  }
  std::vector<uint64_t> words(total, 0);
  for (size_t i = 0; i < nq; i++) {
    if (!pack_query_string(s + offsets[i], slens[i], words.data() + word_off[i])) {
      set_error("query %zu holds a byte other than A/C/G/T: the reference compares raw bytes there, which the 2-bit index "
                "cannot reproduce (rejected, not answered differently)", i);
      return -1;
    }
  }
  uint64_t *d_words = nullptr, *d_off = nullptr;
  uint32_t *d_slen = nullptr, *d_len = nullptr;
  long long *d_km = nullptr, *d_o = nullptr;
  int rc = -1;
  do {
    if (cudaMalloc(&d_words, total * 8) || cudaMalloc(&d_off, nq * 8) || cudaMalloc(&d_slen, nq * 4) ||
        cudaMalloc(&d_km, nq * 8) || cudaMalloc(&d_o, nq * 8) || (lengths && cudaMalloc(&d_len, nq * 4))) {
      set_error("query_str_batch: device allocation failed");
      break;
    }
    cudaMemcpy(d_words, words.data(), total * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_off, word_off.data(), nq * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_slen, slens, nq * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_km, kmers, nq * 8, cudaMemcpyHostToDevice);
    if (lengths) cudaMemcpy(d_len, lengths, nq * 4, cudaMemcpyHostToDevice);
    if (launch_string_query(ix->view(), d_words, d_off, d_slen, d_len, d_km, nq, d_o, 0)) break;
    cudaError_t e = cudaMemcpy(out, d_o, nq * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("query_str_batch: %s", cudaGetErrorString(e)); break; }
    rc = 0;
  } while (0);
  cudaFree(d_words); cudaFree(d_off); cudaFree(d_slen); cudaFree(d_len); cudaFree(d_km); cudaFree(d_o);
  return rc;
}

int64_t sapling_b200_query_str(sapling_b200_index* ix, const char* s, size_t slen, int64_t kmer, size_t length) {
  if (!ix) { set_error("null index"); return -2; }
  if (length > slen) { set_error("plQuery: length > s.length() reads past the string in the reference"); return -2; }
  if (slen <= sapling_b200_index::kSingleMaxBases) {
    std::lock_guard<std::mutex> lock(ix->mu1);
    cudaSetDevice(ix->device);
    if (!ix->m1) {
      if (cudaStreamCreateWithFlags(&ix->s1, cudaStreamNonBlocking) != cudaSuccess ||
          cudaHostAlloc(reinterpret_cast<void**>(&ix->m1), 72 * 8, cudaHostAllocMapped) != cudaSuccess) {
        set_error("plQuery: cannot allocate the mapped staging block");
        return -2;
      }
    }
    uint64_t* m = ix->m1;
    const size_t nw = (slen + 31) / 32 + 1;
    for (size_t i = 0; i < nw; i++) m[i] = 0;
    if (!pack_query_string(s, slen, m)) {
      set_error("plQuery: the query holds a byte other than A/C/G/T: the reference compares raw bytes there, which the "
                "2-bit index cannot reproduce (rejected, not answered differently)");
      return -2;
    }
    m[66] = 0;
    m[67] = (uint64_t)kmer;
    m[68] = (uint64_t)(int64_t)-2;
    uint32_t* sl = reinterpret_cast<uint32_t*>(m + 69);
    sl[0] = (uint32_t)slen;
    sl[1] = (uint32_t)length;
    if (launch_string_query(ix->view(), m, m + 66, sl, sl + 1, reinterpret_cast<const long long*>(m + 67), 1,
                            reinterpret_cast<long long*>(m + 68), ix->s1))
      return -2;
    cudaError_t e = cudaStreamSynchronize(ix->s1);
    if (e != cudaSuccess) { set_error("plQuery: %s", cudaGetErrorString(e)); return -2; }
    return (int64_t)m[68];
  }
  const uint64_t off = 0;
  const uint32_t sl = (uint32_t)slen, ln = (uint32_t)length;
  int64_t out = -1;
  if (sapling_b200_query_str_batch(ix, s, &off, &sl, &ln, &kmer, 1, &out)) return -2;
  return out;
}

int sapling_b200_predict_batch(sapling_b200_index* ix, const uint64_t* kmers, size_t nq, uint64_t* out) {
  if (!ix) { set_error("null index"); return -1; }
  if (nq == 0) return 0;
  cudaSetDevice(ix->device);
  uint64_t *d_k = nullptr, *d_o = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_k, nq * 8));
  if (cudaMalloc(&d_o, nq * 8) != cudaSuccess) { cudaFree(d_k); set_error("predict_batch: allocation failed"); return -1; }
  cudaMemcpy(d_k, kmers, nq * 8, cudaMemcpyHostToDevice);
  int rc = launch_predict(ix->view(), d_k, nq, d_o, 0);
  cudaError_t e = cudaMemcpy(out, d_o, nq * 8, cudaMemcpyDeviceToHost);
  cudaFree(d_k);
  cudaFree(d_o);
  if (rc) return rc;
  SB_CUDA_CHECK(e);
  return 0;
}

int sapling_b200_count_hits(sapling_b200_index* ix, const uint32_t* sa_pos, size_t count, uint32_t maxHits,
                            uint32_t* left, uint32_t* right) {
  if (!ix) { set_error("null index"); return -1; }
  if (!ix->d_kflag) { set_error("count_hits: k-prefix flags not resident (open with SAPLING_B200_KEEP_BUILD)"); return -1; }
  if (count == 0) return 0;
  cudaSetDevice(ix->device);
  uint32_t *d_p = nullptr, *d_l = nullptr, *d_r = nullptr;
  if (cudaMalloc(&d_p, count * 4) || cudaMalloc(&d_l, count * 4) || cudaMalloc(&d_r, count * 4)) {
    cudaFree(d_p); cudaFree(d_l); cudaFree(d_r);
    set_error("count_hits: allocation failed");
    return -1;
  }
  cudaMemcpy(d_p, sa_pos, count * 4, cudaMemcpyHostToDevice);
  int rc = count_hits(ix->d_kflag, ix->n, ix->k, d_p, count, maxHits, d_l, d_r, 0);
  cudaMemcpy(left, d_l, count * 4, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(right, d_r, count * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_p); cudaFree(d_l); cudaFree(d_r);
  if (rc) return rc;
  SB_CUDA_CHECK(e);
  return 0;
}

static int seed_args_ok(sapling_b200_index* ix, uint32_t num_seeds, uint32_t max_hits) {
  if (!ix) { set_error("null index"); return -1; }
  if (!ix->d_kflag || !ix->d_isa) {
    set_error("seed_batch: inverse suffix array / k-prefix flags not resident (open with SAPLING_B200_KEEP_BUILD)");
    return -1;
  }
  if (num_seeds == 0) { set_error("seed_batch: num_seeds must be >= 1"); return -1; }
  if (max_hits > 255) { set_error("seed_batch: max_hits must be <= 255 (hit counts travel as bytes)"); return -1; }
  return 0;
}

// Blocks of reads flow through the same three-stream pipeline as the k-mer batches: the upload of block i+1 and the
// download of block i-1 overlap the seed kernel of block i.  Per seed 10 bytes come back (ref_pos 4, sa_pos 4, left 1,
// right 1).
int sapling_b200_seed_batch_compact(sapling_b200_index* ix, const char* reads, const uint64_t* read_off, size_t n_reads,
                                    uint32_t num_seeds, uint32_t max_hits, uint32_t* ref_pos, uint32_t* sa_pos,
                                    uint8_t* left, uint8_t* right) {
  if (seed_args_ok(ix, num_seeds, max_hits)) return -1;
  if (n_reads == 0) return 0;
  std::lock_guard<std::mutex> lock(ix->mu);
  cudaSetDevice(ix->device);
  if (ensure_staging(ix)) return -1;
  using Slot = sapling_b200_index::SeedSlot;
  constexpr int NS = sapling_b200_index::kSeedSlots;
  const size_t per_read = 2 * (size_t)num_seeds;
  const size_t RB = std::max<size_t>(1, std::min<size_t>(n_reads, ((size_t)1 << 20) / per_read));  // reads per block
  const size_t nblocks = (n_reads + RB - 1) / RB;
  // pinned caller buffers are copied from / to directly; pageable ones go through the slots' pinned mirrors
  const bool pin_in = is_pinned(reads);
  const bool pin_out = is_pinned(ref_pos) && is_pinned(sa_pos) && is_pinned(left) && is_pinned(right);
  auto grow = [&](Slot& s, size_t read_bytes, size_t seeds) -> int {
    if (s.off_cap < RB + 1) {
      cudaFree(s.d_off);
      if (s.h_off) cudaFreeHost(s.h_off);
      s.d_off = nullptr; s.h_off = nullptr; s.off_cap = 0;
      if (cudaMalloc(&s.d_off, (RB + 1) * 8) != cudaSuccess || cudaMallocHost(&s.h_off, (RB + 1) * 8) != cudaSuccess) return -1;
      s.off_cap = RB + 1;
    }
    if (s.reads_cap < read_bytes) {
      const size_t cap = read_bytes + (read_bytes >> 2) + 64;
      cudaFree(s.d_reads);
      if (s.h_reads) cudaFreeHost(s.h_reads);
      s.d_reads = nullptr; s.h_reads = nullptr; s.reads_cap = 0;
      if (cudaMalloc(&s.d_reads, cap) != cudaSuccess || cudaMallocHost(&s.h_reads, cap) != cudaSuccess) return -1;
      s.reads_cap = cap;
    }
    if (s.seeds_cap < seeds) {
      cudaFree(s.d_rp); cudaFree(s.d_sp); cudaFree(s.d_l); cudaFree(s.d_r);
      if (s.h_rp) cudaFreeHost(s.h_rp);
      if (s.h_sp) cudaFreeHost(s.h_sp);
      if (s.h_l) cudaFreeHost(s.h_l);
      if (s.h_r) cudaFreeHost(s.h_r);
      s.d_rp = s.d_sp = s.h_rp = s.h_sp = nullptr; s.d_l = s.d_r = s.h_l = s.h_r = nullptr; s.seeds_cap = 0;
      if (cudaMalloc(&s.d_rp, seeds * 4) != cudaSuccess || cudaMalloc(&s.d_sp, seeds * 4) != cudaSuccess ||
          cudaMalloc(&s.d_l, seeds) != cudaSuccess || cudaMalloc(&s.d_r, seeds) != cudaSuccess ||
          cudaMallocHost(&s.h_rp, seeds * 4) != cudaSuccess || cudaMallocHost(&s.h_sp, seeds * 4) != cudaSuccess ||
          cudaMallocHost(&s.h_l, seeds) != cudaSuccess || cudaMallocHost(&s.h_r, seeds) != cudaSuccess)
        return -1;
      s.seeds_cap = seeds;
    }
    return 0;
  };
  cudaStream_t s_up = ix->streams[0], s_k = ix->streams[1], s_down = ix->streams[2];
  const IndexView v = ix->view();
  int rc = 0;
  for (size_t b = 0; b < nblocks + NS && !rc; b++) {
    if (b >= (size_t)NS) {  // retire block b - NS: its slot is about to be reused
      const size_t rb = b - NS;
      Slot& s = ix->seed_slots[rb % NS];
      if (cudaEventSynchronize(ix->ev_down[rb % NS]) != cudaSuccess) { rc = -1; break; }
      if (!pin_out) {
        const size_t r0 = rb * RB, m = std::min(RB, n_reads - r0), t0 = r0 * per_read, tm = m * per_read;
        memcpy(ref_pos + t0, s.h_rp, tm * 4);
        memcpy(sa_pos + t0, s.h_sp, tm * 4);
        memcpy(left + t0, s.h_l, tm);
        memcpy(right + t0, s.h_r, tm);
      }
    }
    if (b < nblocks) {
      Slot& s = ix->seed_slots[b % NS];
      const int e = (int)(b % NS);
      const size_t r0 = b * RB, m = std::min(RB, n_reads - r0);
      const uint64_t base = read_off[r0], nbytes = read_off[r0 + m] - base;
      if (grow(s, (size_t)nbytes, RB * per_read)) { cudaGetLastError(); set_error("seed_batch: allocation failed"); rc = -2; break; }
      for (size_t i = 0; i <= m; i++) s.h_off[i] = read_off[r0 + i] - base;
      const char* src = reads + base;
      if (!pin_in) {
        memcpy(s.h_reads, src, (size_t)nbytes);
        src = s.h_reads;
      }
      if (cudaMemcpyAsync(s.d_reads, src, nbytes, cudaMemcpyHostToDevice, s_up) != cudaSuccess ||
          cudaMemcpyAsync(s.d_off, s.h_off, (m + 1) * 8, cudaMemcpyHostToDevice, s_up) != cudaSuccess) { rc = -1; break; }
      cudaEventRecord(ix->ev_up[e], s_up);
      cudaStreamWaitEvent(s_k, ix->ev_up[e], 0);
      if (launch_seeds(v, ix->d_isa, ix->d_kflag, s.d_reads, s.d_off, m, num_seeds, max_hits, s.d_rp, s.d_sp, s.d_l, s.d_r, s_k)) { rc = -2; break; }
      cudaEventRecord(ix->ev_k[e], s_k);
      cudaStreamWaitEvent(s_down, ix->ev_k[e], 0);
      const size_t t0 = r0 * per_read, tm = m * per_read;
      cudaMemcpyAsync(pin_out ? ref_pos + t0 : s.h_rp, s.d_rp, tm * 4, cudaMemcpyDeviceToHost, s_down);
      cudaMemcpyAsync(pin_out ? sa_pos + t0 : s.h_sp, s.d_sp, tm * 4, cudaMemcpyDeviceToHost, s_down);
      cudaMemcpyAsync(pin_out ? left + t0 : s.h_l, s.d_l, tm, cudaMemcpyDeviceToHost, s_down);
      if (cudaMemcpyAsync(pin_out ? right + t0 : s.h_r, s.d_r, tm, cudaMemcpyDeviceToHost, s_down) != cudaSuccess) { rc = -1; break; }
      cudaEventRecord(ix->ev_down[e], s_down);
    }
  }
  if (rc) {
    for (int i = 0; i < 3; i++) cudaStreamSynchronize(ix->streams[i]);  // nothing in flight behind an error
    if (rc == -1) set_error("seed_batch: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return 0;  // every block was retired (event-synchronised) inside the loop
}

// The same tuples in the reference-shaped types (int64 positions with -1, 32-bit counts).
int sapling_b200_seed_batch(sapling_b200_index* ix, const char* reads, const uint64_t* read_off, size_t n_reads,
                            uint32_t num_seeds, uint32_t max_hits, int64_t* ref_pos, uint32_t* sa_pos, uint32_t* left,
                            uint32_t* right) {
  if (seed_args_ok(ix, num_seeds, max_hits)) return -1;
  const size_t total = n_reads * 2 * (size_t)num_seeds;
  std::vector<uint32_t> rp(total);
  std::vector<uint8_t> l(total), r(total);
  if (sapling_b200_seed_batch_compact(ix, reads, read_off, n_reads, num_seeds, max_hits, rp.data(), sa_pos, l.data(), r.data()))
    return -1;
  for (size_t i = 0; i < total; i++) {
    ref_pos[i] = rp[i] == 0xFFFFFFFFu ? -1 : (int64_t)rp[i];
    left[i] = l[i];
    right[i] = r[i];
  }
  return 0;
}

int sapling_b200_seed_batch_dev(sapling_b200_index* ix, const char* d_reads, const uint64_t* d_read_off, size_t n_reads,
                                uint32_t num_seeds, uint32_t max_hits, uint32_t* d_ref_pos, uint32_t* d_sa_pos,
                                uint8_t* d_left, uint8_t* d_right, void* stream) {
  if (seed_args_ok(ix, num_seeds, max_hits)) return -1;
  cudaSetDevice(ix->device);
  return launch_seeds(ix->view(), ix->d_isa, ix->d_kflag, d_reads, d_read_off, n_reads, num_seeds, max_hits, d_ref_pos,
                      d_sa_pos, d_left, d_right, static_cast<cudaStream_t>(stream));
}

uint64_t sapling_b200_oob_count(sapling_b200_index* ix) {
  if (!ix || !ix->d_oob) return 0;
  unsigned long long total = 0;
  auto one = [&](sapling_b200_index* d) {
    unsigned long long v = 0;
    cudaSetDevice(d->device);
    cudaMemcpy(&v, d->d_oob, 8, cudaMemcpyDeviceToHost);
    total += v;
  };
  one(ix);
  for (auto* r : ix->replicas) one(r);
  cudaSetDevice(ix->device);
  return total;
}

int sapling_b200_sample_queries_dev(sapling_b200_index* ix, uint64_t seed, uint64_t mut_seed, uint64_t first,
                                    size_t nq, uint64_t* d_kmers, void* stream) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  return launch_sample(ix->view(), seed, mut_seed, first, nq, d_kmers, static_cast<cudaStream_t>(stream));
}

int sapling_b200_verify_dev(sapling_b200_index* ix, const uint64_t* d_kmers, const int64_t* d_out, size_t nq,
                            uint64_t* n_match, uint64_t* n_minus1, void* stream) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* d_c = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_c, 16));
  SB_CUDA_CHECK(cudaMemsetAsync(d_c, 0, 16, st));
  int rc = launch_verify(ix->view(), d_kmers, reinterpret_cast<const long long*>(d_out), nq, d_c, st);
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d_c, 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_c);
  if (rc) return rc;
  SB_CUDA_CHECK(e);
  if (n_match) *n_match = h[0];
  if (n_minus1) *n_minus1 = h[1];
  return 0;
}

int sapling_b200_count_probes_dev(sapling_b200_index* ix, const uint64_t* d_kmers, size_t nq, uint64_t* n_probes,
                                  void* stream) {
  if (!ix) { set_error("null index"); return -1; }
  cudaSetDevice(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* d_c = nullptr;
  SB_CUDA_CHECK(cudaMalloc(&d_c, 8));
  SB_CUDA_CHECK(cudaMemsetAsync(d_c, 0, 8, st));
  int rc = launch_probe_count(ix->view(), d_kmers, nq, d_c, st);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, d_c, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_c);
  if (rc) return rc;
  SB_CUDA_CHECK(e);
  if (n_probes) *n_probes = h;
  return 0;
}

int sapling_b200_gather_bench(uint64_t bytes, uint64_t n_loads, int reps, double* gbps) {
  int dev = 0;
  if (require_device(&dev)) return -1;
  return run_gather_bench(bytes, n_loads, reps, gbps);
}

int sapling_b200_gather_bench2(uint64_t bytes, uint64_t n_access, int gran, int chain, int blocks_per_sm, int reps,
                               double* gacc_per_s) {
  int dev = 0;
  if (require_device(&dev)) return -1;
  if (gran != 16 && gran != 32 && gran != 64 && gran != 128) { set_error("gran must be 16, 32, 64 or 128"); return -1; }
  return run_gather_bench2(bytes, n_access, gran, chain, blocks_per_sm, reps, gacc_per_s);
}

}  // extern "C"

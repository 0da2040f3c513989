/*
 * sapling_example_batched.cpp -- the reference's benchmark driver (mkirsche/sapling src/sapling_example.cpp) with its
 * query loop handed to the GPU a batch at a time.
 *
 *   sapling_example_batched <genome.fa> [saFn=..] [sapFn=..] [errFn=..] [nb=..] [k=..] [nq=..] [maxMem=..] [qLen=..]
 *
 * Same command line (sapling_example.cpp:43-84), same experiments (k-10, k, k+10, k+20, k+30, k+80, or qLen; :91-103), same
 * queries (the reference samples them with unseeded rand(), :115 -- the same sequence here), the same queries.out side
 * file (:120-129) and the same two report lines per experiment ("Piecewise linear time: ..", "Piecewise linear
 * correctness: X out of N", :141,:154).  The unmodified driver also runs on the drop-in header (INTEGRATION.md section 1;
 * the drivers_b200 build), but it calls plQuery once per query -- a kernel launch and a host round trip each; this one times what the
 * library is built for: ONE call per experiment,
 *   query length == k : Sapling::queryBatch (k-mers as 64-bit words; sharded over the GPUs of SAPLING_B200_GPUS)
 *   any other length  : Sapling::plQueryBatch (strings; the gallop loops of sapling_api.h:184-196,:229-241 on the GPU)
 * with host buffers in and out, copies inside the timed region like the reference's loop.
 */
#include <chrono>

#include "sapling_api.h" /* include/sapling_api.h: the drop-in struct Sapling */

static int k = -1, numBuckets = -1, maxMem = -1, numQueries = 5000000, queryLength = -1; /* sapling_example.cpp:18-24 */
static Sapling sap;

static void run_experiment(int qlen) {
  cout << "Running experiment to search for " << qlen << "-mers" << endl;
  vector<string> queries(numQueries, "");
  vector<long long> kmers(numQueries, 0);
  for (int i = 0; i < numQueries; i++) {
    const size_t idx = rand() % (sap.n - qlen); /* :115 */
    queries[i] = sap.reference.substr(idx, qlen);
    kmers[i] = sap.kmerizeAdjusted(qlen, queries[i]);
  }
  FILE *outfile = fopen("queries.out", "w"); /* :120-129 */
  for (int i = 0; i < numQueries; i++) {
    fprintf(outfile, "@read%d\n%s\n+\n", i + 1, queries[i].c_str());
    for (int j = 0; j < qlen; j++) fputc('9', outfile);
    fputc('\n', outfile);
  }
  fclose(outfile);
  cout << "Constructed queries" << endl;

  vector<long long> plAnswers(numQueries, 0);
  auto start = std::chrono::system_clock::now();
  if (qlen == sap.k) {
    sap.queryBatch(reinterpret_cast<const uint64_t *>(kmers.data()), (size_t)numQueries, plAnswers.data());
  } else {
    plAnswers = sap.plQueryBatch(queries, kmers, vector<size_t>((size_t)numQueries, (size_t)qlen));
  }
  auto end = std::chrono::system_clock::now();
  std::chrono::duration<double> elapsed_seconds = end - start;
  cout << "Piecewise linear time: " << elapsed_seconds.count() << endl;

  int countCorrect = 0; /* :144-154 */
  for (int i = 0; i < numQueries; i++) {
    if (plAnswers[i] == -1) continue;
    if (plAnswers[i] + (long long)qlen <= (long long)sap.n && queries[i] == sap.reference.substr(plAnswers[i], qlen))
      countCorrect++;
  }
  cout << "Piecewise linear correctness: " << countCorrect << " out of " << numQueries << endl;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    cout << "Usage: " << argv[0] << " <genome> [saFn=<suffix array file>] [sapFn=<sapling file>] [nb=<log number of buckets>] "
         << "[maxMem=<max number of buckets will be (genome size)/val>] [k=<k>] [nq=<number of queries>] "
         << "[errFn=<errors file if outputting them>] [qLen=<query length>]" << endl;
    return 0;
  }
  string refFnString = argv[1];
  string saFnString = refFnString + ".sa", saplingFnString = refFnString + ".sap", errorFnString = "";
  for (int i = 2; i < argc; i++) {
    string cur = argv[i];
    size_t eqPos = cur.find("=");
    if (eqPos == string::npos) continue;
    string arg = cur.substr(0, eqPos), val = cur.substr(eqPos + 1);
    if (arg == "saFn") saFnString = val;
    if (arg == "sapFn") saplingFnString = val;
    if (arg == "errFn") errorFnString = val;
    if (arg == "nb") numBuckets = stoi(val);
    if (arg == "k") k = stoi(val);
    if (arg == "nq") numQueries = stoi(val);
    if (arg == "maxMem") maxMem = stoi(val);
    if (arg == "qLen") queryLength = stoi(val);
  }
  sap = Sapling(refFnString, saFnString, saplingFnString, numBuckets, maxMem, k, errorFnString);
  cout << "Testing Sapling" << endl;
  if (queryLength == -1) {
    run_experiment(sap.k - 10);
    run_experiment(sap.k);
    run_experiment(sap.k + 10);
    run_experiment(sap.k + 20);
    run_experiment(sap.k + 30);
    run_experiment(sap.k + 80);
  } else {
    run_experiment(queryLength);
  }
  return 0;
}

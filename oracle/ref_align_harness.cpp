/*
 * ref_align_harness.cpp -- runs the UNMODIFIED reference aligner (src/align.cpp) with the one thing its snapshot
 * forgets: Sapling::sa is declared (sapling_api.h:38) and read (align.cpp:287-289) but never filled, so the shipped
 * `align` segfaults after "Aligning reads" (SURVEY section 0.1).  The array its author meant is the inverse suffix array
 * (lsa.inv: text position -> rank).
 *
 * TEST INFRASTRUCTURE ONLY.  No SAPLING or aligner logic lives here: align.cpp is #included where it lies (its main
 * renamed, its class members made reachable), and this file's main repeats align.cpp:391-402 with the one assignment
 * added.  Output: oracle/_ref/align_ref (git-ignored).  Used by tests/test_gpu_drivers.py as the SAM oracle for
 * sapling_b200/host/align_b200.cpp.
 */
/* every standard header the included sources pull in, first, so that the keyword games below cannot touch them */
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <tuple>
#include <vector>

#include "ssw_cpp.h" /* has an include guard: align.cpp's own #include becomes a no-op */

#define main reference_align_main
#define class struct /* SaplingAligner keeps its Sapling* private (align.cpp:151-157) */
#include "align.cpp"
#undef class
#undef main

int main(int argc, char **argv)
{
  if (argc < 4)
  {
    usage();
    return 1;
  }
  parseArgs(argc, argv);
  SaplingAligner *al = new SaplingAligner(queryFn, refFn);
  al->sapling->sa = al->sapling->lsa.inv; /* the missing assignment */
  al->align_all_reads(outFn, argc, argv);
  fflush(NULL); /* align_all_reads returns without fclose (align.cpp:219) */
  delete al;
  return 0;
}

#!/usr/bin/env python
"""One small batch through every batch path (partitioned at several slice counts, in-order unpartitioned, one query per
thread; k <= and > the lines' prefix; int64 / uint32 / bit-stream entry points), for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_paths.py
  compute-sanitizer --tool racecheck python tools/sanitize_paths.py
Answers are compared with the one-query-per-thread kernel's.  Nothing here is on the product path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import sapling_b200 as S
    import _fixtures as F
    from sapling_b200.api import pack_kmer_bits
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 60001
    for name, k in (("rand200k", 21), ("gc1991", 31), ("tandem50", 16)):
        g = F.small_genomes()[name]
        kmers = np.tile(F.query_mix(g, k, 30000, seed=3), (nq + 29999) // 30000)[:nq]
        os.environ["SAPLING_B200_TUNE"] = "part=0,inorder_min=-1"
        a = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET)
        exp = a.queryBatch(kmers)
        a.close()
        for tune in ("part=0,inorder_min=1", "part_min=1,part_bits=3,chunk_log2=22", "part_min=1,part_bits=9,chunk_log2=22",
                     "part_min=1,part_bits=11,chunk_log2=16"):
            os.environ["SAPLING_B200_TUNE"] = tune
            b = S.Sapling.from_memory(g, None, k=k, flags=S.QUIET)
            ok = bool(np.array_equal(b.queryBatch(kmers), exp))
            g32 = b.queryBatchBits(pack_kmer_bits(kmers, 2 * k), 2 * k, len(kmers))
            ok32 = bool(np.array_equal(np.where(g32 == 0xFFFFFFFF, -1, g32.astype(np.int64)), exp))
            print(name, k, tune, b.query_kernel(len(kmers))[0], "int64", ok, "bits/u32", ok32, flush=True)
            b.close()
            assert ok and ok32


if __name__ == "__main__":
    main()

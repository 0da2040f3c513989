"""sapling_b200 -- B200 (sm_100a) implementation of SAPLING's suffix-array query hot path.

The product is the C-ABI shared library ``libsapling_b200.so`` (sources in ``csrc/``, header in
``include/sapling_b200.h``) plus the C++ drop-in ``include/sapling_api.h``.  This package is the
thin ctypes binding the tests and ``bench.py`` use; it mirrors the reference's ``struct Sapling``
surface (same member / method names and argument meaning).  There is no CPU fallback.
"""
from .api import (Sapling, SaplingError, kmerize, kmerize_adjusted, lib, lib_path, gather_bench,  # noqa: F401
                  gather_bench2, pack_kmer_bits)

QUIET = 1
NO_COMPAT = 2
KEEP_BUILD = 4

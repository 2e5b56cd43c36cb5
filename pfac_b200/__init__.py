"""pfac_b200 -- B200-native (sm_100a) implementation of PFAC's failureless Aho-Corasick path.

The product is the C-ABI shared library pfac_b200/lib/libpfac.so (include/PFAC.h is the
reference's public interface; include/PFAC_ext.h the additive 64-bit / shard / table-compiler
entry points).  This Python package is a thin ctypes mirror of that ABI for tests, bench.py
and tooling; it contains no matching logic and no CPU fallback.
"""
from .api import (PFAC, PFACComm, PFACError, PFACMultiGPU, TableCompiler, Status, Platform, PerfMode, TextureMode,  # noqa: F401
                  load_library, library_path)

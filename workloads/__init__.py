"""Synthetic workloads of the BASELINE.json configs -- test and bench
tooling, not part of the product (pfac_b200/ holds only the matching path).

  synth   numpy generators: counter-hash text, pattern sets, planting (the definition)
  devgen  the same text + planting regenerated on the GPU by a small CUDA library
          (workloads/csrc/devgen.cu -> workloads/lib/libpfac_devgen.so), byte-equal to synth
"""
from . import synth  # noqa: F401

#!/usr/bin/env python
"""Bandwidth of the library's host zero fill (PFAC_hostZero: the copy pool + streaming stores), which
bounds the sparse result path of PFAC_matchFromHost at 4 bytes per input byte.  CPU only.

    [PFAC_B200_COPY_THREADS=n] python tools/host_zero_bw.py
"""
import numpy as np, time, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfac_b200.api import load_library
L = load_library()
n = 1 << 32
dst = np.ones(n, np.uint8)
best = 0
for rep in range(4):
    t = time.perf_counter(); L.PFAC_hostZero(dst.ctypes.data, n); dt = time.perf_counter() - t
    best = max(best, n / dt / 1e9)
print('threads', os.environ.get('PFAC_B200_COPY_THREADS', 'default'), 'zero GB/s', round(best, 1), 'cpus', os.cpu_count())

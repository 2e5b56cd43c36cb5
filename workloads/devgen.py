"""ctypes front of workloads/lib/libpfac_devgen.so: synth.make_text on the GPU (torch tensors for
the device memory).  Byte-equal to the numpy definition in synth.py (tests/test_devgen.py)."""
import ctypes
import os

import numpy as np

from . import build as _build
from . import synth

_KIND = {"random": 0, "ascii": 1, "dna": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            _build.build()
        L = ctypes.CDLL(_build.LIB)
        u64 = ctypes.c_ulonglong
        L.pfac_devgen_text.argtypes = [ctypes.c_int, u64, u64, u64, ctypes.c_void_p, ctypes.c_void_p]
        L.pfac_devgen_plant.argtypes = [ctypes.c_void_p, u64, u64, u64, ctypes.c_void_p, ctypes.c_void_p, u64, u64,
                                        u64, u64, u64, u64, ctypes.c_void_p]
        _lib = L
    return _lib


class DevicePatterns:
    """Pattern bytes + offsets resident on one device (for planting)."""

    def __init__(self, patterns, device):
        import torch
        lens = np.array([len(p) for p in patterns], dtype=np.uint64)
        off = np.zeros(len(patterns) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        blob = np.frombuffer(b"".join(patterns) or b"\0", dtype=np.uint8)
        self.n = len(patterns)
        self.maxlen = int(lens.max()) if len(patterns) else 0
        self.longest = int(np.argmax(lens)) if len(patterns) else 0   # first of maximum length, as max(key=len)
        self.bytes = torch.from_numpy(blob.copy()).to(device)
        self.off = torch.from_numpy(off.view(np.int64).copy()).to(device)


def make_text(kind, seed, start, n, total_len, patterns=None, every=4096, device="cuda", out=None,
              boundary=1 << 20):
    """synth.make_text(kind, seed, start, n, total_len, patterns, every) as a uint8 tensor on `device`.
    `patterns` may be a list of bytes or a DevicePatterns (reuse it across calls)."""
    import torch
    dev = torch.device(device)
    if out is None:
        out = torch.empty(n, dtype=torch.uint8, device=dev)
    assert out.is_cuda and out.dtype == torch.uint8 and out.numel() >= n and out.is_contiguous()
    stream = torch.cuda.current_stream(dev).cuda_stream
    L = lib()
    with torch.cuda.device(dev):
        rc = L.pfac_devgen_text(_KIND[kind], seed & synth.MASK64, start, n, out.data_ptr(), stream)
        if rc:
            raise RuntimeError("pfac_devgen_text failed: %d" % rc)
        if patterns is not None and every:
            dp = patterns if isinstance(patterns, DevicePatterns) else DevicePatterns(patterns, dev)
            if dp.n:
                rc = L.pfac_devgen_plant(out.data_ptr(), start, n, total_len, dp.bytes.data_ptr(), dp.off.data_ptr(),
                                         dp.n, dp.maxlen, dp.longest, seed & synth.MASK64, every, boundary, stream)
                if rc:
                    raise RuntimeError("pfac_devgen_plant failed: %d" % rc)
    return out[:n]

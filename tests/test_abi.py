"""The C-ABI library loads and exports every symbol include/*.h declares; status strings and
argument checks that need no GPU behave like the reference (PFAC.cpp).  CPU only: no compute."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from pfac_b200 import PFAC, PFACError, Status, load_library, library_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(PFAC_[A-Za-z0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = load_library()
    names = _declared("PFAC.h") + _declared("PFAC_ext.h")
    assert len(_declared("PFAC.h")) == 12  # the reference's 12 entry points (PFAC.h:87-215)
    for n in names:
        assert hasattr(L, n), "libpfac.so does not export " + n
    out = subprocess.run(["nm", "-D", "--defined-only", library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (PFAC_\w+)", out))
    assert set(names) <= exported
    # no CPU matcher and nothing from oracle/ in the product library
    assert not any(s.startswith(("orc_", "ref_")) for s in re.findall(r" T (\w+)", out))


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_enum_values_match_reference_abi():
    assert Status.SUCCESS == 0 and Status.BASE == 10000
    assert [int(s) for s in (Status.ALLOC_FAILED, Status.CUDA_ALLOC_FAILED, Status.INVALID_HANDLE,
                             Status.INVALID_PARAMETER, Status.PATTERNS_NOT_READY, Status.FILE_OPEN_ERROR,
                             Status.LIB_NOT_EXIST, Status.ARCH_MISMATCH, Status.MUTEX_ERROR,
                             Status.INTERNAL_ERROR)] == list(range(10001, 10011))
    hdr = open(os.path.join(ROOT, "include", "PFAC.h")).read()
    for frag in ["PFAC_PLATFORM_GPU = 0", "PFAC_PLATFORM_CPU = 1", "PFAC_PLATFORM_CPU_OMP = 2",
                 "PFAC_AUTOMATIC = 0", "PFAC_TEXTURE_ON = 1", "PFAC_TEXTURE_OFF = 2",
                 "PFAC_TIME_DRIVEN = 0", "PFAC_SPACE_DRIVEN = 1", "PFAC_STATUS_BASE = 10000"]:
        assert frag in hdr


def test_error_strings():
    L = load_library()
    assert L.PFAC_getErrorString(0) == b"PFAC_STATUS_SUCCESS: operation is successful"
    assert L.PFAC_getErrorString(10003) == b"PFAC_STATUS_INVALID_HANDLE: handle is invalid (NULL)"
    assert L.PFAC_getErrorString(10005).startswith(b"PFAC_STATUS_PATTERNS_NOT_READY")
    assert L.PFAC_getErrorString(10010) == b"PFAC_STATUS_INTERNAL_ERROR: please report bugs"
    assert L.PFAC_getErrorString(2) == b"out of memory"  # < BASE: forwarded to cudaGetErrorString


def test_null_handle_is_invalid_handle_first():
    """Reference order: NULL handle -> INVALID_HANDLE before anything else (PFAC.cpp:846,882,967)."""
    L = load_library()
    n = ctypes.c_int(0)
    assert L.PFAC_destroy(None) == Status.INVALID_HANDLE
    assert L.PFAC_setPlatform(None, 0) == Status.INVALID_HANDLE
    assert L.PFAC_setTextureMode(None, 0) == Status.INVALID_HANDLE
    assert L.PFAC_setPerfMode(None, 0) == Status.INVALID_HANDLE
    assert L.PFAC_readPatternFromFile(None, b"x") == Status.INVALID_HANDLE
    assert L.PFAC_matchFromDevice(None, None, 0, None) == Status.INVALID_HANDLE
    assert L.PFAC_matchFromHost(None, None, 0, None) == Status.INVALID_HANDLE
    assert L.PFAC_matchFromDeviceReduce(None, None, 0, None, None, ctypes.byref(n)) == Status.INVALID_HANDLE
    assert L.PFAC_matchFromHostReduce(None, None, 0, None, None, ctypes.byref(n)) == Status.INVALID_HANDLE
    assert L.PFAC_dumpTransitionTable(None, None) == Status.INVALID_HANDLE
    # additive entry points (include/PFAC_ext.h) follow the same rule
    m = ctypes.c_ulonglong(0)
    assert L.PFAC_reduceOnDevice(None, None, 0, None, None, ctypes.byref(n)) == Status.INVALID_HANDLE
    assert L.PFAC_reduceInplaceOnDevice(None, None, 0, None, None, ctypes.byref(n)) == Status.INVALID_HANDLE
    assert L.PFAC_setStream(None, None) == Status.INVALID_HANDLE
    assert L.PFAC_readPatternFromMemory(None, b"x\n", 2) == Status.INVALID_HANDLE
    assert L.PFAC_readPatternFromArrays(None, None, None, 0) == Status.INVALID_HANDLE
    assert L.PFAC_matchShardFromDevice(None, None, 0, 0, None) == Status.INVALID_HANDLE
    assert L.PFAC_matchFromDeviceReduce64(None, None, 0, None, None, ctypes.byref(m)) == Status.INVALID_HANDLE
    assert L.PFAC_matchShardFromDeviceReduce64(None, None, 0, 0, 0, None, None, ctypes.byref(m)) == Status.INVALID_HANDLE
    assert L.PFAC_getTableInfo(None, None) == Status.INVALID_HANDLE
    assert L.PFAC_getTableInfoReduce(None, None) == Status.INVALID_HANDLE
    assert L.PFAC_memoryUsage(None) == Status.INVALID_HANDLE
    assert L.PFAC_dumpTransitionTableToFile(None, b"x") == Status.INVALID_HANDLE
    assert L.PFAC_saveCompiledPatterns(None, b"x") == Status.INVALID_HANDLE
    assert L.PFAC_loadCompiledPatterns(None, b"x") == Status.INVALID_HANDLE
    assert L.PFAC_lastHostTransfer(None, None, None) == Status.INVALID_HANDLE
    assert L.PFAC_tableSave(None, b"x") == Status.INVALID_HANDLE
    assert L.PFAC_tableDestroy(None) == Status.INVALID_HANDLE
    assert L.PFAC_mgpuDestroy(None) == Status.INVALID_HANDLE
    assert L.PFAC_mgpuMatchFromHost(None, None, 0, None) == Status.INVALID_HANDLE


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a device PFAC_create returns the raw CUDA error (reference
    PFAC.cpp:148-151 does the same)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(PFACError) as e:
        PFAC()
    assert 0 < e.value.status < Status.BASE


def test_host_copy_pool_is_exact_and_thread_safe():
    """PFAC_hostCopy = the copy pool behind the pageable-buffer pipelines (pfac_api.cu CopyPool): odd
    sizes and alignments, zero bytes, and several caller threads sharing the one pool."""
    import threading
    from pfac_b200.api import host_copy, load_library
    L = load_library()
    assert L.PFAC_hostCopy(None, None, 0) == 0
    assert L.PFAC_hostCopy(None, None, 8) == Status.INVALID_PARAMETER
    rng = np.random.default_rng(9)
    for n in (1, 63, 64, 65, (1 << 20) - 1, (1 << 20) + 1, 9_437_189):
        src = rng.integers(0, 256, size=n + 3, dtype=np.uint8)[3:]      # misaligned source
        dst = np.zeros(n + 5, dtype=np.uint8)
        host_copy(dst[5:], src)
        assert np.array_equal(dst[5:], src) and not dst[:5].any()
    for n in (1, 63, 65, 70_001, (1 << 20) + 3, 5_000_011):                   # PFAC_hostZero: the pool's zero fill
        buf = np.full(n + 9, 7, dtype=np.uint8)
        assert L.PFAC_hostZero(buf[4:].ctypes.data, n) == 0
        assert not buf[4:4 + n].any() and (buf[:4] == 7).all() and (buf[4 + n:] == 7).all()
    srcs = [rng.integers(0, 256, size=6_000_000 + 4097 * i, dtype=np.uint8) for i in range(4)]
    dsts = [np.zeros_like(x) for x in srcs]
    ts = [threading.Thread(target=lambda d=d, x=x: [host_copy(d, x) for _ in range(3)]) for d, x in zip(dsts, srcs)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(np.array_equal(d, x) for d, x in zip(dsts, srcs))


def test_headers_are_plain_c_and_the_example_links(tmp_path):
    """The boundary is a C ABI: include/PFAC.h + PFAC_ext.h compile as C89/C99 and as C++, and a C caller
    links against libpfac.so with nothing but -lpfac (reference PFAC/test/Makefile:23)."""
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    src = tmp_path / "caller.c"
    src.write_text('#include "PFAC.h"\n#include "PFAC_ext.h"\n#include <stdio.h>\n'
                   "int main(void) {\n"
                   "    PFAC_tableInfo_t info; (void)info;\n"
                   "    if (PFAC_destroy(0) != PFAC_STATUS_INVALID_HANDLE) return 1;\n"
                   '    printf("%s\\n", PFAC_getErrorString(PFAC_STATUS_INVALID_HANDLE));\n'
                   "    return 0;\n}\n")
    for cc, std in (("/usr/bin/gcc", "-std=c89"), ("/usr/bin/gcc", "-std=c99"), ("/usr/bin/g++", "-std=c++11")):
        extra = ["-x", "c++"] if cc.endswith("g++") else []
        subprocess.run([cc, std, "-Wall", "-Werror", "-fsyntax-only", "-I", inc] + extra + [str(src)], check=True)
    exe = tmp_path / "caller"
    libdir = os.path.dirname(library_path())
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", inc, str(src), "-L", libdir, "-lpfac",
                    "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("PFAC_STATUS_INVALID_HANDLE")


def test_multi_gpu_example_compiles_and_links(tmp_path):
    """examples/multi_gpu_scan.cpp (the omp_PFAC.cpp shape on PFAC_comm) is written against the public
    headers only: it must compile with warnings as errors and link with -lpfac -lcudart."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    libdir = os.path.dirname(library_path())
    subprocess.run(["/usr/bin/g++", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", cuda + "/include",
                    os.path.join(root, "examples", "multi_gpu_scan.cpp"), "-L", libdir, "-lpfac", "-L", cuda + "/lib64",
                    "-lcudart", "-Wl,-rpath," + libdir, "-o", str(tmp_path / "multi_gpu_scan")], check=True)

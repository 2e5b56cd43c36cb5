"""Builds pfac_b200/lib/libpfac.so (the C-ABI product library) in-tree with nvcc for sm_100a.

    python -m pfac_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libpfac.so")
SOURCES = ["pfac_api.cu", "pfac_kernels.cu", "pfac_table.cpp"]
HEADERS = ["pfac_kernels.h", "pfac_table.h"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; use the system g++
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(ROOT, "include", f) for f in ("PFAC.h", "PFAC_ext.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [
        _nvcc(), "-ccbin", _host_cxx(), "-std=c++17", "-O3",
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
        "-o", LIB,
    ] + list(extra_flags) + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libpfac.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Builds workloads/lib/libpfac_devgen.so (device-side workload generator) with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "devgen.cu")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libpfac_devgen.so")


def needs_build():
    return not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB)


def build(force=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [nvcc, "-ccbin", cxx, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", LIB, SRC]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
        raise RuntimeError("nvcc failed building libpfac_devgen.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

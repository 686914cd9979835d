"""Builds libtinyimgcodec_cuda.so in-tree with nvcc for sm_100a (no other target)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtinyimgcodec_cuda.so")
SOURCES = ["tic_encode.cu"]
HEADERS = ["tic_kernels.cuh", "tic_tables.h", os.path.join("..", "..", "include", "tinyimgcodec_cuda.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)

"""Builds libtinyimgcodec_cuda.so in-tree with nvcc for sm_100a (no other target).

One object per translation unit under csrc/_obj/ (only stale ones are recompiled), then one link."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libtinyimgcodec_cuda.so")
PUBLIC = os.path.join("..", "..", "include", "tinyimgcodec_cuda.h")
# translation unit -> the headers it includes
SOURCES = {
    "tic_encode.cu": ["tic_kernels.cuh", "tic_tables.h", PUBLIC],
    "tic_decode.cu": ["tic_tables.h", PUBLIC],
}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _obj(src):
    return os.path.join(OBJ, src[:-3] + ".o")


def needs_build():
    return any(_stale(_obj(s), [os.path.join(CSRC, f) for f in [s] + hs]) for s, hs in SOURCES.items()) or \
        _stale(LIB, [_obj(s) for s in SOURCES if os.path.exists(_obj(s))])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-warn-spills"]

    def compile_one(src):
        deps = [os.path.join(CSRC, f) for f in [src] + SOURCES[src]]
        if not force and not _stale(_obj(src), deps):
            return None
        return subprocess.run(common + ["-c", "-o", _obj(src), os.path.join(CSRC, src)],
                              capture_output=True, text=True)

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for res in results:
        if res is None:
            continue
        if verbose or res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode:
            raise RuntimeError("nvcc failed")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] +
                         [_obj(s) for s in SOURCES], capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)

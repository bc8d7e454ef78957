"""Build libsnacc_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNACC_B200_LIB") or os.path.join(HERE, "libsnacc_b200.so")   # override: experiments only
SOURCES = [os.path.join(HERE, "csrc", "api.cu")]
# every header api.cu can include: all of csrc/*.cuh plus the public C header (a stale-check that misses one -- the
# packed LZ4 kernels live in lz4_packed.cuh / pack.cuh -- lets tests run against an old binary)
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + sorted(
    glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsnacc_b200.so cannot be built")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    subprocess.check_call(cmd)
    return LIB_PATH

"""CPU: the product's __host__ __device__ parse functions (snacc_b200/csrc/*.cuh), compiled for the host
by tests/host_emu.cu, against the oracle.  This checks the exact kernel logic -- two-segment stream
accessor, prefix checkpoint, block budget rules -- without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import lib
from oracle.make_golden import VECTOR_KINDS, synth_vector
from snacc_b200 import _build

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "libhost_emu.so")
    src = os.path.join(HERE, "host_emu.cu")
    deps = [src] + _build.HEADERS
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([_build._nvcc(), "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                               "-shared", "-o", so, src])
    e = ctypes.CDLL(so)
    e.emu_lz4_size.restype = ctypes.c_int64
    e.emu_lz4_size.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
    return e


def lz4_emu(emu, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.uint8)
    if y is None:
        return emu.emu_lz4_size(x.ctypes.data, x.size, None, -1)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    return emu.emu_lz4_size(x.ctypes.data, x.size, y.ctypes.data, y.size)


@pytest.mark.parametrize("kind", VECTOR_KINDS)
def test_lz4_kernel_logic_singles_and_pairs(emu, kind):
    rng = np.random.default_rng(11)
    for lx in [1, 12, 13, 700, 11000, 65535, 65536, 65537, 70000, 131072, 150000]:
        x = synth_vector(kind, lx, lx + 1)
        assert lz4_emu(emu, x) == lib.lz4f_size(x), (kind, lx)
        for ly in [1, 13, 9000, 65536, 80000]:
            y = synth_vector(kind if rng.random() < 0.7 else "dna", ly, ly + 7)
            assert lz4_emu(emu, x, y) == lib.lz4f_size(np.concatenate([x, y])), (kind, lx, ly)


def test_lz4_related_genomes_cross_boundary_matches(emu):
    from snacc_b200 import synth
    g = synth.phylogeny(4, 40000, seed=3)
    for a in g:
        for b in g:
            assert lz4_emu(emu, a, b) == lib.ref_lz4f_size(np.concatenate([a, b]))

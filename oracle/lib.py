"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# codec ids shared with include/snacc_b200.h
CODEC_IDS = {"lz4": 0, "gzip": 1, "zlib": 2}
# bytes the reference's compressor call adds around the raw deflate stream
# (gzip.compress: 10-byte header + 8-byte trailer; zlib.compress: 2 + 4)
CODEC_WRAPPER_BYTES = {"lz4": 0, "gzip": 18, "zlib": 6}


def build():
    """Compile liboracle.so with gcc (build the checker; not the product)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    u8p = ctypes.c_void_p
    lib.oracle_lz4f_size.restype = ctypes.c_uint64
    lib.oracle_lz4f_size.argtypes = [u8p, ctypes.c_uint64]
    lib.oracle_lz4f_size_ex.restype = ctypes.c_uint64
    lib.oracle_lz4f_size_ex.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32]
    lib.oracle_deflate_size.restype = ctypes.c_uint64
    lib.oracle_deflate_size.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int]
    lib.oracle_deflate_bits.restype = ctypes.c_uint64
    lib.oracle_deflate_bits.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.ref_lz4f_size.restype = ctypes.c_int64
    lib.ref_lz4f_size.argtypes = [u8p, ctypes.c_uint64]
    lib.ref_lz4f_size_flag.restype = ctypes.c_int64
    lib.ref_lz4f_size_flag.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int]
    lib.ref_deflate_size.restype = ctypes.c_int64
    lib.ref_deflate_size.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int]
    lib.ref_compressed_len.restype = ctypes.c_int64
    lib.ref_compressed_len.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int]
    lib.ref_batch_sizes.restype = ctypes.c_int
    lib.ref_batch_sizes.argtypes = [u8p, u8p, u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
    lib.ref_lz4_version.restype = ctypes.c_int
    lib.ref_zlib_version.restype = ctypes.c_char_p
    _LIB = lib
    return lib


def _buf(data):
    """bytes / numpy uint8 -> (padded numpy array keeping 16 readable slack bytes, length)."""
    a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.asarray(data, dtype=np.uint8)
    n = a.size
    p = np.zeros(n + 16, dtype=np.uint8)
    p[:n] = a
    return p, n


def lz4f_size(data, content_size_flag=1):
    """len(lz4framed.compress(data)) by the C restatement."""
    p, n = _buf(data)
    return int(load().oracle_lz4f_size_ex(p.ctypes.data, n, content_size_flag, None, 0))


def lz4f_block_sizes(data):
    p, n = _buf(data)
    nb = max(1, (n + 65535) // 65536)
    out = np.zeros(nb, dtype=np.uint32)
    load().oracle_lz4f_size_ex(p.ctypes.data, n, 1, out.ctypes.data, nb)
    return out


def deflate_size(data, level):
    """len(raw deflate stream) by the C restatement (level 9 or 6, memLevel 8, default strategy)."""
    p, n = _buf(data)
    return int(load().oracle_deflate_size(p.ctypes.data, n, level))


def deflate_stats(data, level):
    p, n = _buf(data)
    nb = ctypes.c_uint64(0)
    ns = ctypes.c_uint64(0)
    bits = int(load().oracle_deflate_bits(p.ctypes.data, n, level, ctypes.byref(nb), ctypes.byref(ns)))
    return bits, nb.value, ns.value


def compressed_len(data, algorithm):
    """len(<reference compressor call>(data)) by the restatements (no +33)."""
    if algorithm == "lz4":
        return lz4f_size(data)
    if algorithm == "gzip":
        return deflate_size(data, 9) + 18
    if algorithm == "zlib":
        return deflate_size(data, 6) + 6
    raise KeyError(algorithm)


def ref_lz4f_size(data, content_size_flag=1):
    p, n = _buf(data)
    r = int(load().ref_lz4f_size_flag(p.ctypes.data, n, content_size_flag))
    if r < 0:
        raise RuntimeError(f"liblz4 call failed: {r}")
    return r


def ref_deflate_size(data, level):
    p, n = _buf(data)
    r = int(load().ref_deflate_size(p.ctypes.data, n, level))
    if r < 0:
        raise RuntimeError(f"zlib call failed: {r}")
    return r


def ref_compressed_len(data, algorithm):
    p, n = _buf(data)
    r = int(load().ref_compressed_len(p.ctypes.data, n, CODEC_IDS[algorithm]))
    if r < 0:
        raise RuntimeError(f"codec call failed: {r}")
    return r


def ref_batch_sizes(corpus, offsets, job_x, job_y, algorithm, n_threads):
    """Threaded batch over pre-loaded sequences using the real system codecs (CPU baseline)."""
    corpus = np.ascontiguousarray(corpus, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    job_x = np.ascontiguousarray(job_x, dtype=np.int32)
    job_y = np.ascontiguousarray(job_y, dtype=np.int32)
    out = np.zeros(job_x.size, dtype=np.int64)
    load().ref_batch_sizes(corpus.ctypes.data, offsets.ctypes.data, job_x.ctypes.data, job_y.ctypes.data,
                           int(job_x.size), CODEC_IDS[algorithm], int(n_threads), out.ctypes.data)
    return out

"""Import the UNMODIFIED reference module from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference/snacc/pairwise_ncd.py`` imports two packages that are
not installed in this image (``lz4framed`` at :12, ``Bio`` at :13).  Two tiny stand-ins are put into
``sys.modules`` so that the reference file itself runs untouched:

* ``lz4framed.compress(b)``  -> ``LZ4F_compressFrame`` of the system liblz4.so.1 (1.9.4) with the
  preferences py-lz4framed's one-shot ``compress`` uses by default (64 KiB blocks, linked, no
  checksums, level 0, contentSize = len(b)) -- see SURVEY.md section 8c.
* ``Bio.SeqIO.parse(path, "fasta")`` -> ``oracle.fasta_shim`` (multi-record FASTA, whitespace
  stripped, case kept; ``Seq.reverse_complement`` with Biopython's ambiguous-DNA table).

Nothing on the GPU box may call this: /root/reference does not exist there.  It is used by
``oracle/make_golden.py`` to produce ``tests/golden/*.json`` and by CPU tests that skip when the
reference tree is absent.
"""
import ctypes
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "snacc", "pairwise_ncd.py"))


class _Prefs(ctypes.Structure):
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int),
                ("contentChecksumFlag", ctypes.c_int), ("frameType", ctypes.c_int),
                ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint),
                ("blockChecksumFlag", ctypes.c_int), ("compressionLevel", ctypes.c_int),
                ("autoFlush", ctypes.c_uint), ("favorDecSpeed", ctypes.c_uint),
                ("reserved", ctypes.c_uint * 3)]


def _make_lz4framed():
    lib = ctypes.CDLL("liblz4.so.1")
    lib.LZ4F_compressFrameBound.restype = ctypes.c_size_t
    lib.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    lib.LZ4F_compressFrame.restype = ctypes.c_size_t
    lib.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p,
                                       ctypes.c_size_t, ctypes.c_void_p]
    lib.LZ4F_isError.restype = ctypes.c_uint
    lib.LZ4F_isError.argtypes = [ctypes.c_size_t]
    lib.LZ4_versionString.restype = ctypes.c_char_p

    def compress(b):
        prefs = _Prefs()
        prefs.contentSize = len(b)
        cap = lib.LZ4F_compressFrameBound(len(b), ctypes.byref(prefs))
        dst = ctypes.create_string_buffer(cap)
        n = lib.LZ4F_compressFrame(dst, cap, b, len(b), ctypes.byref(prefs))
        if lib.LZ4F_isError(n):
            raise RuntimeError("LZ4F_compressFrame failed")
        return dst.raw[:n]

    mod = types.ModuleType("lz4framed")
    mod.compress = compress
    mod.__version__ = "shim(liblz4 %s)" % lib.LZ4_versionString().decode()
    return mod


def _make_bio():
    from . import fasta_shim
    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse = fasta_shim.parse
    bio.SeqIO = seqio
    return bio, seqio


def install_shims():
    if "lz4framed" not in sys.modules:
        sys.modules["lz4framed"] = _make_lz4framed()
    if "Bio" not in sys.modules:
        bio, seqio = _make_bio()
        sys.modules["Bio"] = bio
        sys.modules["Bio.SeqIO"] = seqio


def load_reference_pairwise_ncd():
    """Return the reference's own ``snacc.pairwise_ncd`` module object, unmodified."""
    if not available():
        raise FileNotFoundError("reference tree not present (expected only in the build container)")
    install_shims()
    spec = importlib.util.spec_from_file_location(
        "_reference_snacc_pairwise_ncd", os.path.join(REFERENCE_ROOT, "snacc", "pairwise_ncd.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_package():
    """Import the reference ``snacc`` package (cli included) under the name ``_reference_snacc``."""
    if not available():
        raise FileNotFoundError("reference tree not present (expected only in the build container)")
    install_shims()
    name = "_reference_snacc"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REFERENCE_ROOT, "snacc", "__init__.py"),
        submodule_search_locations=[os.path.join(REFERENCE_ROOT, "snacc")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod

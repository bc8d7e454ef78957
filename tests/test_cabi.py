"""CPU: libsnacc_b200.so loads and exports every symbol include/snacc_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from snacc_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "snacc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snacc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("snacc_ctx_create", "snacc_ctx_destroy", "snacc_upload", "snacc_upload_device", "snacc_single_sizes",
              "snacc_pair_sizes", "snacc_tile_sizes", "snacc_ncd", "snacc_last_error", "snacc_version"):
        assert s in syms


def test_library_builds_for_sm100a_and_exports_all_symbols():
    _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/snacc_b200.h but not exported"
    lib.snacc_version.restype = ctypes.c_int
    assert lib.snacc_version() >= 100


def test_sass_is_sm100a_only():
    import shutil
    import subprocess
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cu, "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(7|8|9)\d", out)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "snacc_b200")
    for dirpath, _, names in os.walk(pkg):
        for nm in names:
            if nm.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, nm)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, nm


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from snacc_b200.engine import Engine, SnaccGpuError
    with pytest.raises(SnaccGpuError):
        Engine(0)

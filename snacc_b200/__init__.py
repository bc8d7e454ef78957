"""snacc_b200 -- B200-native replacement for the all-pairs NCD hot path of alexsweeten/snacc.

Public surface mirrors the reference package (snacc/__init__.py:1-2): ``compressed_size``,
``compute_distance``, ``__version__``; plus ``ncd_matrix`` (the batch entry point) and ``Engine``
(the ctypes binding of libsnacc_b200.so).
"""
__version__ = "0.1.0"

from .pairwise_ncd import compressed_size, compute_distance, extract_sequences, ncd_matrix  # noqa: E402,F401
from .engine import Engine, SnaccGpuError  # noqa: E402,F401

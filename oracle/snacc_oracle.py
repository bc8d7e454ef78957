"""CPU restatement of the reference's NCD path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows, function by function:
  extract_sequences   /root/reference/snacc/pairwise_ncd.py:15-39
  compressed_size     /root/reference/snacc/pairwise_ncd.py:42-90   (size = len(compressed) + 33)
  compute_distance    /root/reference/snacc/pairwise_ncd.py:93-111
  all-pairs driver    /root/reference/snacc/cli.py:104-142          (N singles, N*N ordered pairs,
                                                                     pivot, CSV)
``backend="port"`` uses the C restatements (lz4_oracle.c / deflate_oracle.c); ``backend="system"``
uses the real system codecs through ref_codecs.c.
"""
import sys
from pathlib import Path

import numpy as np

from . import fasta_shim, lib

GETSIZEOF_BIAS = sys.getsizeof(b"")   # 33 on 64-bit CPython 3.x (pairwise_ncd.py:90)
SUPPORTED = ("lz4", "gzip", "zlib")


def extract_sequences(sequences, reverse_complement=False):
    """pairwise_ncd.py:15-39 -- concatenated record sequences; tuple -> extract(a) + extract(b)."""
    if type(sequences) == tuple:
        return (extract_sequences(sequences[0], reverse_complement)
                + extract_sequences(sequences[1], reverse_complement))
    parts = []
    for rec in fasta_shim.parse(Path(sequences).absolute(), "fasta"):
        parts.append(str(rec.seq.reverse_complement()) if reverse_complement else str(rec.seq))
    seq = "".join(parts)
    if not seq:
        raise ValueError(f"No sequence extracted. Ensure that file {Path(sequences).absolute()} contains a "
                         "proper FASTA definition line (i.e. a line that starts with '>sequence_name').")
    return seq


def compressed_len(data, algorithm, backend="port"):
    if algorithm in ("lzma", "bzip2"):
        # not on the GPU path yet: the reference's own calls, pairwise_ncd.py:71-76 -- the oracle a
        # device implementation will be held to (golden: tests/golden/reference_sizes_lzma_bzip2.json)
        import bz2
        import lzma
        return len(lzma.compress(bytes(data)) if algorithm == "lzma" else bz2.compress(bytes(data)))
    if algorithm not in SUPPORTED:
        raise KeyError(algorithm)
    if backend == "port":
        return lib.compressed_len(data, algorithm)
    return lib.ref_compressed_len(data, algorithm)


def compressed_size(sequences, algorithm, reverse_complement=False, backend="port"):
    """pairwise_ncd.py:42-90 -- returns (sequences, len(compressed) + 33)."""
    data = extract_sequences(sequences, reverse_complement).encode("utf-8")
    return sequences, compressed_len(data, algorithm, backend) + GETSIZEOF_BIAS


def compute_distance(x, y, cxy, cyx):
    """pairwise_ncd.py:93-111 -- min over both concatenation orders of (C - min(x,y)) / max(x,y)."""
    lo, hi = (y, x) if x > y else (x, y)
    return min((cxy - lo) / hi, (cyx - lo) / hi)


def size_tables(files, algorithm, reverse_complement=False, backend="port"):
    """cli.py:108-129 -- C[i] for every file and S[i][j] for every ORDERED pair incl. the diagonal.
    Sizes are raw compressed lengths (no +33).  Each file is parsed once (the reference re-parses
    per job; the result is the same)."""
    seqs = [extract_sequences(Path(f), reverse_complement).encode("utf-8") for f in files]
    n = len(seqs)
    C = np.zeros(n, dtype=np.int64)
    S = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        C[i] = compressed_len(seqs[i], algorithm, backend)
        for j in range(n):
            S[i, j] = compressed_len(seqs[i] + seqs[j], algorithm, backend)
    return C, S


def ncd_from_sizes(C, S, bias=GETSIZEOF_BIAS):
    """cli.py:131-136 -- D[i][j] = compute_distance(C_i+33, C_j+33, S_ij+33, S_ji+33) in float64."""
    n = len(C)
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            D[i, j] = compute_distance(int(C[i]) + bias, int(C[j]) + bias,
                                       int(S[i, j]) + bias, int(S[j, i]) + bias)
    return D


def sorted_files(paths):
    """cli.py:89-102 -- files + suffix-filtered directory contents, de-duplicated, sorted by str(abs)."""
    paths = [Path(p) for p in paths]
    files = [p for p in paths if p.is_file()]
    for d in [p for p in paths if p.is_dir()]:
        for f in d.iterdir():
            if f.suffix.lower() in [".fasta", ".fna", ".fa", ".faa", ".fsa"]:
                files.append(f)
    return sorted(set(files), key=lambda p: str(p.absolute()))


def write_csv(files, D, output):
    """cli.py:138-142 -- long table -> pivot(index='file', columns='file2') -> to_csv."""
    import pandas as pd
    rows = [(files[i], files[j], D[i, j]) for i in range(len(files)) for j in range(len(files))]
    df = pd.DataFrame(rows, columns=["file", "file2", "ncd"])
    df = df.pivot(index="file", columns="file2", values="ncd")
    df.to_csv(output)

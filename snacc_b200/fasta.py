"""FASTA loading for the GPU path: each file is parsed ONCE (the reference re-parses every file for
every job, snacc/pairwise_ncd.py:29-36) into one byte string plus its record lengths.

Behaviour mirrors what the reference gets from ``Bio.SeqIO.parse(path, "fasta")``: records start at
'>' lines, anything before the first '>' is ignored, sequence lines are right-stripped and joined,
blanks and carriage returns inside a record are dropped, case is preserved.  Reverse complement is NOT
done here: record boundaries are handed to the device, which reverse-complements each record
(snacc_upload, K0).
"""
import os
from pathlib import Path

import numpy as np

_STRIP = bytes([9, 10, 11, 12, 13, 32])




def _read_fasta_lines(raw):
    """line-by-line parser: the reference semantics spelled out (text mode = universal newlines, rstrip of every
    line, then blanks dropped)"""
    import re
    recs = []
    cur = None
    for line in re.split(rb"\r\n|\r|\n", raw):
        if line.startswith(b">"):
            if cur is not None:
                recs.append(b"".join(cur))
            cur = []
        elif cur is not None:
            cur.append(line.rstrip(_STRIP).replace(b" ", b"").replace(b"\r", b""))
    if cur is not None:
        recs.append(b"".join(cur))
    return recs


def _read_fasta_native(raw):
    """snacc_fasta_parse of libsnacc_b200.so: the same line semantics in C (a host function of the C ABI; the call
    releases the GIL, so load_corpus parses files on several threads)"""
    import ctypes
    from .engine import load_library
    lib = load_library()
    n = len(raw)
    out = np.empty(max(n, 1), dtype=np.uint8)
    out_len = ctypes.c_uint64(0)
    cap = 64
    while True:
        rec_len = np.zeros(cap, dtype=np.uint64)
        src = ctypes.cast(ctypes.c_char_p(raw), ctypes.c_void_p)         # the bytes object's own buffer: read-only use
        recs = int(lib.snacc_fasta_parse(src, n, out.ctypes.data, ctypes.byref(out_len), rec_len.ctypes.data, cap))
        if recs < 0:
            raise RuntimeError(f"snacc_fasta_parse failed ({recs})")
        if recs <= cap:
            return out[:out_len.value], [int(v) for v in rec_len[:recs]]
        cap = recs                                           # more records than expected: once more with room for all


def read_fasta(path):
    """-> (uint8 array of all records' residues concatenated in file order, list of record lengths)."""
    raw = Path(path).read_bytes()
    from .engine import SnaccGpuError
    try:
        return _read_fasta_native(raw)
    except (SnaccGpuError, OSError, AttributeError):         # library not built yet: the same host-side semantics in
        recs = _read_fasta_lines(raw)                        # Python (parsing is host work either way; every
        return np.frombuffer(b"".join(recs), dtype=np.uint8), [len(r) for r in recs]   # compressor call still needs the GPU)


def load_corpus(files):
    """-> (data uint8, seq_offsets uint64[n+1], rec_offsets uint64[m+1]).  Raises the reference's
    ValueError for a file without any sequence (pairwise_ncd.py:37-38)."""
    datas, seq_off, rec_off = [], [0], [0]
    files = list(files)
    if len(files) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, len(files), os.cpu_count() or 1)) as pool:
            parsed = list(pool.map(read_fasta, files))       # file reads and the native parser release the GIL
    else:
        parsed = [read_fasta(f) for f in files]
    for f, (d, recs) in zip(files, parsed):
        if d.size == 0:
            raise ValueError(f"No sequence extracted. Ensure that file {Path(f).absolute()} contains a proper FASTA "
                             "definition line (i.e. a line that starts with '>sequence_name').")
        datas.append(d)
        seq_off.append(seq_off[-1] + d.size)
        for r in recs:
            rec_off.append(rec_off[-1] + r)
    data = np.concatenate(datas) if datas else np.zeros(0, np.uint8)
    return data, np.asarray(seq_off, dtype=np.uint64), np.asarray(rec_off, dtype=np.uint64)

"""FASTA loading for the GPU path: each file is parsed ONCE (the reference re-parses every file for
every job, snacc/pairwise_ncd.py:29-36) into one byte string plus its record lengths.

Behaviour mirrors what the reference gets from ``Bio.SeqIO.parse(path, "fasta")``: records start at
'>' lines, anything before the first '>' is ignored, sequence lines are right-stripped and joined,
blanks and carriage returns inside a record are dropped, case is preserved.  Reverse complement is NOT
done here: record boundaries are handed to the device, which reverse-complements each record
(snacc_upload, K0).
"""
from pathlib import Path

import numpy as np

_STRIP = bytes([9, 10, 11, 12, 13, 32])


_DROP = b" \r\n"


def _read_fasta_lines(raw):
    """line-by-line parser: the reference semantics spelled out (rstrip of every line, then blanks and CRs dropped)"""
    recs = []
    cur = None
    for line in raw.split(b"\n"):
        if line.startswith(b">"):
            if cur is not None:
                recs.append(b"".join(cur))
            cur = []
        elif cur is not None:
            cur.append(line.rstrip(_STRIP).replace(b" ", b"").replace(b"\r", b""))
    if cur is not None:
        recs.append(b"".join(cur))
    return recs


def read_fasta(path):
    """-> (uint8 array of all records' residues concatenated in file order, list of record lengths).

    Files without tabs, vertical tabs or form feeds (all real FASTA files) take a path that never splits into lines:
    the records are cut at the '>' line starts and each body is filtered in one ``bytes.translate`` -- rstrip of
    every line followed by dropping blanks and CRs is then the same as deleting blanks, CRs and newlines.  c5 is
    2 048 files of 5 MB: seconds instead of a minute of Python line handling."""
    raw = Path(path).read_bytes()
    if b"\t" in raw or b"\x0b" in raw or b"\x0c" in raw:
        recs = _read_fasta_lines(raw)
    else:
        starts = [0] if raw.startswith(b">") else []          # '>' at a line start (bytes.find: a MULTILINE regex
        i = raw.find(b"\n>")                                  # over 70 000 lines costs more than the whole parse)
        while i >= 0:
            starts.append(i + 1)
            i = raw.find(b"\n>", i + 1)
        recs = []
        for k, a in enumerate(starts):
            b = starts[k + 1] if k + 1 < len(starts) else len(raw)
            eol = raw.find(b"\n", a, b)
            recs.append(raw[eol + 1:b].translate(None, _DROP) if eol >= 0 else b"")
    lengths = [len(r) for r in recs]
    data = np.frombuffer(b"".join(recs), dtype=np.uint8)
    return data, lengths


def load_corpus(files):
    """-> (data uint8, seq_offsets uint64[n+1], rec_offsets uint64[m+1]).  Raises the reference's
    ValueError for a file without any sequence (pairwise_ncd.py:37-38)."""
    datas, seq_off, rec_off = [], [0], [0]
    for f in files:
        d, recs = read_fasta(f)
        if d.size == 0:
            raise ValueError(f"No sequence extracted. Ensure that file {Path(f).absolute()} contains a proper FASTA "
                             "definition line (i.e. a line that starts with '>sequence_name').")
        datas.append(d)
        seq_off.append(seq_off[-1] + d.size)
        for r in recs:
            rec_off.append(rec_off[-1] + r)
    data = np.concatenate(datas) if datas else np.zeros(0, np.uint8)
    return data, np.asarray(seq_off, dtype=np.uint64), np.asarray(rec_off, dtype=np.uint64)

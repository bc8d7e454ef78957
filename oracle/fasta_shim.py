"""Minimal stand-in for ``Bio.SeqIO.parse(path, "fasta")`` (TEST INFRASTRUCTURE ONLY).

Restates the behaviour the reference relies on at /root/reference/snacc/pairwise_ncd.py:32-36:
records start at lines beginning with '>', text before the first '>' is ignored, sequence lines are
right-stripped and joined, spaces and carriage returns inside them are dropped, case is preserved.
``record.seq.reverse_complement()`` uses Biopython's ambiguous-DNA complement table (plus U -> A),
case-preserving; characters outside the table are left unchanged.
"""

_PAIRS = {"A": "T", "C": "G", "G": "C", "T": "A", "M": "K", "R": "Y", "W": "W", "S": "S", "Y": "R",
          "K": "M", "V": "B", "H": "D", "D": "H", "B": "V", "X": "X", "N": "N", "U": "A"}
_TABLE = bytes(range(256))
_tab = bytearray(_TABLE)
for _k, _v in _PAIRS.items():
    _tab[ord(_k)] = ord(_v)
    _tab[ord(_k.lower())] = ord(_v.lower())
COMPLEMENT_TABLE = bytes(_tab)


class _Seq:
    def __init__(self, s):
        self._s = s

    def __str__(self):
        return self._s

    def __len__(self):
        return len(self._s)

    def reverse_complement(self):
        return _Seq(self._s.encode("latin-1").translate(COMPLEMENT_TABLE)[::-1].decode("latin-1"))


class _Record:
    def __init__(self, title, seq):
        self.id = title.split(None, 1)[0] if title.split() else ""
        self.description = title
        self.seq = _Seq(seq)


def parse(path, fmt="fasta"):
    if fmt != "fasta":
        raise ValueError("only fasta is supported by the shim")
    title = None
    lines = []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    yield _Record(title, "".join(lines).replace(" ", "").replace("\r", ""))
                title = line[1:].rstrip()
                lines = []
            elif title is not None:
                lines.append(line.rstrip())
    if title is not None:
        yield _Record(title, "".join(lines).replace(" ", "").replace("\r", ""))

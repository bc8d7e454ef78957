"""oracle/ -- CPU checker for the snacc all-pairs NCD hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product package ``snacc_b200`` never does.

Contents
  * ``lz4_oracle.c`` / ``deflate_oracle.c`` : plain-C restatements of the compressed-size functions
    behind ``/root/reference/snacc/pairwise_ncd.py:74,78,80``.
  * ``ref_codecs.c`` : wrappers over the real system codecs (liblz4 1.9.4, zlib 1.3) those calls
    resolve to in this image -- the thing the restatements are pinned against.
  * ``snacc_oracle.py`` : restatement of ``pairwise_ncd.py`` / ``cli.py:104-142`` semantics
    (FASTA extraction, +33 ``sys.getsizeof`` bias, NCD formula, CSV) on top of the above.
  * ``ref_loader.py`` : imports the UNMODIFIED reference module from /root/reference with two
    ``sys.modules`` shims (``lz4framed`` -> liblz4 via ctypes, ``Bio.SeqIO`` -> minimal parser);
    only usable in the build container, used to generate ``tests/golden``.

Parity pinning status: the reference holds no golden vector for any lz4/gzip compressed size
(SURVEY.md 8c).  The oracle is pinned against outputs of the reference itself run in this container
(``tests/golden/*.json`` produced by ``oracle/make_golden.py``) and against the system codecs.
"""
from .lib import (load, lz4f_size, deflate_size, compressed_len, ref_lz4f_size, ref_deflate_size,
                  ref_compressed_len, ref_batch_sizes, CODEC_IDS, CODEC_WRAPPER_BYTES)

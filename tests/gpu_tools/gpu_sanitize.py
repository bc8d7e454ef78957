"""small lz4 / gzip / zlib jobs for compute-sanitizer (memcheck / racecheck): every kernel of the library runs once"""
import sys
import numpy as np
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine
from oracle import lib as olib
g = synth.phylogeny(3, 90_000, seed=11) + synth.phylogeny(2, 3_000, seed=12)
n = len(g)
with Engine(0) as eng:
    eng.upload_sequences(g, reverse_complement=False)
    for algo in ("lz4", "gzip", "zlib"):
        C = eng.single_sizes(algo)
        S = eng.tile_sizes(algo, 0, n, 0, n)
        ref = np.array([[olib.ref_compressed_len(np.concatenate([a, b]), algo) for b in g] for a in g])
        assert np.array_equal(S, ref), algo
        assert np.array_equal(C, np.array([olib.ref_compressed_len(a, algo) for a in g])), algo
        print(algo, "ok", int(S.sum()))

"""Dev script (GPU box): deflate kernels -- parity against zlib and first timings."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from snacc_b200.engine import Engine
from snacc_b200 import synth
from oracle import lib as olib
eng = Engine(0)
def dna(n, seed): return np.random.default_rng(seed).choice(np.frombuffer(b"ACGT", np.uint8), size=n)
G = int(os.environ.get("G", "8")); L = int(os.environ.get("L", "4500000"))
import torch
big = [g.cpu().numpy() for g in synth.phylogeny_torch(G, L, 2, torch.device("cuda", 0))]
eng.upload_sequences(big)
for algo, lvl in (("gzip", 9), ("zlib", 6)):
    t = time.time(); C = eng.single_sizes(algo); dt = time.time() - t
    print(algo, "singles s", dt, "kernel ms", eng.last_kernel_ms())
    w = 18 if algo == "gzip" else 6
    t = time.time(); ref0 = olib.ref_deflate_size(big[0], lvl) + w; print("cpu one single s", time.time() - t)
    print("check single0", C[0], ref0)
    t = time.time(); S = eng.tile_sizes(algo, 0, G, 0, G); dt = time.time() - t
    print(algo, "pairs s", dt, "kernel ms", eng.last_kernel_ms())
    for a, b in [(0, 1), (G - 1, 2)]:
        print("check pair", a, b, S[a, b], olib.ref_deflate_size(np.concatenate([big[a], big[b]]), lvl) + w)
    t = time.time(); S2 = eng.tile_sizes(algo, 0, G, 0, G); print(algo, "pairs again (caches warm) s", time.time() - t, bool((S == S2).all()))

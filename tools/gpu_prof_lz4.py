"""Dev script (GPU box): a small packed-LZ4 tile run for ncu (G genomes x L bases, all ordered pairs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from snacc_b200.engine import Engine
G = int(os.environ.get("G", "48")); L = int(os.environ.get("L", "1000000"))
rng = np.random.default_rng(7)
seqs = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L + int(rng.integers(-2000, 2000)))] for _ in range(G)]
eng = Engine(0)
eng.upload_sequences(seqs)
t = time.time(); C = eng.single_sizes("lz4"); print("singles", time.time() - t)
t = time.time(); S = eng.tile_sizes("lz4", 0, G, 0, G); print("pairs", time.time() - t, eng.stat("main_kernel_ms"), int(S.sum()))

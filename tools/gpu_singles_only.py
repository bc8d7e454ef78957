"""GPU box: the LZ4 singles + prefix checkpoint pass alone (148 genomes of 5 Mbp), for timing / ncu."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 148, 5_000_000
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(n + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
for rep in range(3):
    eng.set_option("invalidate_caches", 1)
    t0 = time.perf_counter(); C = eng.single_sizes("lz4"); t1 = time.perf_counter()
    print(f"singles of {n} x {L}: {1e3 * (t1 - t0):.1f} ms (kernels {eng.stat('total_kernel_ms'):.1f} ms), checksum {int(C.sum())}", flush=True)

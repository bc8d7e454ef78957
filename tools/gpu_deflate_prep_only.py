"""GPU box: gzip preparation of m sequences of the c4 corpus (for ncu: dfl_prep_kernel and friends)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine

m, L = int(sys.argv[1]) if len(sys.argv) > 1 else 148, 5_000_000
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(m, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(m + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
for rep in range(2):
    eng.set_option("invalidate_caches", 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); C = eng.single_sizes("gzip"); t1 = time.perf_counter()
    print(f"gzip preparation of {m} sequences: {1e3 * (t1 - t0):.0f} ms wall, kernels {eng.stat('total_kernel_ms'):.0f} ms, checksum {int(C.sum())}", flush=True)

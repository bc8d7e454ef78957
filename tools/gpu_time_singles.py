"""time the LZ4 singles/checkpoint pass (lz4_pk_single_kernel) on N x 5 Mbp genomes"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, 5_000_000, 4, dev)
so = np.zeros(n + 1, dtype=np.uint64); so[1:] = np.cumsum([x.numel() for x in g])
corpus = torch.cat(g)
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
for it in range(3):
    eng.set_option("invalidate_caches", 1)
    t = time.perf_counter(); c = eng.single_sizes("lz4"); dt = time.perf_counter() - t
    print("singles pass %d: %.1f ms wall, %.1f ms kernels, checksum %d" % (it, dt * 1e3, eng.stat("total_kernel_ms"), int(c.sum())))

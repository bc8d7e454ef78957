"""GPU box: host-side breakdown of one c4 step (where wall time beyond the kernels goes)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import sharding, synth
from snacc_b200.engine import Engine

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 5_000_000
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(n + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
for rep in range(3):
    t0 = time.perf_counter(); eng.set_option("invalidate_caches", 1)
    t1 = time.perf_counter(); C = eng.single_sizes("lz4"); k1 = eng.stat("total_kernel_ms")
    t2 = time.perf_counter(); S = eng.tile_sizes("lz4", 0, n, 0, n); k2 = eng.stat("total_kernel_ms"); m2 = eng.stat("main_kernel_ms")
    t3 = time.perf_counter(); D = eng.ncd(C, S)
    t4 = time.perf_counter()
    st = {}
    eng.set_option("invalidate_caches", 1)
    C2, S2 = sharding.sizes_matrix(eng, "lz4", False, None, st)
    t5 = time.perf_counter()
    print(f"singles {1e3*(t2-t1):.0f} ms (kernel {k1:.0f}), tiles {1e3*(t3-t2):.0f} ms (kernel {k2:.0f}, pair kernel {m2:.0f}), ncd {1e3*(t4-t3):.0f} ms, "
          f"sizes_matrix {1e3*(t5-t4):.0f} ms (kernel {st['kernel_ms']:.0f})")

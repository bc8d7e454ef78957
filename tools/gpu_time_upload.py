"""GPU box: where an e2e step's upload time goes (c4 shape): H2D + scatter + histogram + pack, and the singles pass."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 5_000_000
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(n + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
host = torch.empty(corpus.numel(), dtype=torch.uint8, pin_memory=True); host.copy_(corpus); torch.cuda.synchronize()
eng = Engine(0)
for rep in range(3):
    t = time.perf_counter(); eng.upload(host.numpy(), so); t1 = time.perf_counter() - t
    t = time.perf_counter(); eng.upload_device(corpus.data_ptr(), so); t2 = time.perf_counter() - t
    t = time.perf_counter(); eng.upload_device(corpus.data_ptr(), so, so, True); t3 = time.perf_counter() - t
    t = time.perf_counter(); C = eng.single_sizes("lz4"); t4 = time.perf_counter() - t
    print(f"upload from pinned host {t1*1e3:.1f} ms, from device {t2*1e3:.1f} ms, from device with reverse complement {t3*1e3:.1f} ms, "
          f"singles pass {t4*1e3:.1f} ms (kernel {eng.stat('total_kernel_ms'):.1f})")

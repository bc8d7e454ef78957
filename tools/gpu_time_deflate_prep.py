"""GPU box: what one rank of an N-GPU gzip run does on the c4 corpus, phase by phase (single GPU stand-in): the
preparation of m = 512 / N sequences, and its band of 512 x (512 / N) junction jobs (which also builds the x-only
3-byte index of the other sequences)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine

n, L = 512, 5_000_000
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(n + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
for m in (512, 64, 8, 1, 64):
    eng.set_option("invalidate_caches", 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); eng.single_sizes("gzip", np.arange(m, dtype=np.int32)); t1 = time.perf_counter()
    k1 = eng.stat("total_kernel_ms")
    print(f"gzip preparation of {m} sequences: {1e3 * (t1 - t0):.0f} ms wall, kernels {k1:.0f} ms", flush=True)
    if m == 64:
        # the other ranks' records would arrive by all-gather; here they are computed (not timed) and kept
        eng.single_sizes("gzip")
        t2 = time.perf_counter(); S = eng.tile_sizes("gzip", 0, n, 0, m); t3 = time.perf_counter()
        print(f"  band {n} x {m}: {1e3 * (t3 - t2):.0f} ms wall, kernels {eng.stat('total_kernel_ms'):.0f} ms, "
              f"main {eng.stat('main_kernel_ms'):.0f} ms", flush=True)

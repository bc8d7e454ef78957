"""GPU box: the linked LZ4 pair kernel with its tiles cut into k segments (option lz4_segments), on the bands one rank
of an N-GPU run works on: c4 corpus (512 x 5 Mbp), columns [0, 512 / N).  Prints the pair-kernel time per k and checks
that every k gives the same sizes."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from snacc_b200 import synth
from snacc_b200.engine import Engine

n, L = 512, 5_000_000
rate = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
dev = torch.device("cuda", 0)
g = synth.phylogeny_torch(n, L, 4, dev)
lengths = np.array([x.numel() for x in g]); so = np.zeros(n + 1, np.uint64); so[1:] = np.cumsum(lengths)
corpus = torch.cat(g); del g
if rate > 0:
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    m = torch.rand(corpus.numel(), device=dev, generator=gen) < rate
    corpus[m] = ord("N")
eng = Engine(0)
eng.upload_device(corpus.data_ptr(), so)
eng.single_sizes("lz4")
for ranks in (8, 4, 1):
    cols = n // ranks
    ref = None
    for k in (1, 2, 4, 8, 0):
        eng.set_option("lz4_segments", k)
        best = 1e9
        for rep in range(2):
            S = eng.tile_sizes("lz4", 0, n, 0, cols)
            best = min(best, eng.stat("main_kernel_ms"))
        used = int(eng.stat("lz4_segments"))
        if ref is None:
            ref = S
        same = bool(np.array_equal(ref, S))
        print(f"ranks {ranks} cols {cols} lz4_segments {k} (used {used}): pair kernel {best:.1f} ms, same sizes {same}, "
              f"{n * cols / best * 1e3:.0f} pairs/s", flush=True)
        assert same
        if ranks == 1 and k == 1:
            pass

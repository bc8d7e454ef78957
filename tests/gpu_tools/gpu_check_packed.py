"""Dev script (GPU box): packed LZ4 tile kernels -- parity against liblz4 and first timings."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from snacc_b200.engine import Engine
from snacc_b200 import synth
from oracle import lib as olib

eng = Engine(0)
rng = np.random.default_rng(5)
def dna(n, seed): return np.random.default_rng(seed).choice(np.frombuffer(b"ACGT", np.uint8), size=n)

# ---- parity: lengths across regimes, all ordered pairs ----
lens = [1, 5, 13, 20, 700, 11000, 30000, 40000, 65520, 65535, 65536, 65537, 70000, 131072, 150000, 300000]
seqs = [dna(n, n) for n in lens] + synth.phylogeny(4, 40000, seed=3) + synth.phylogeny(3, 200000, seed=4)
eng.upload_sequences(seqs)
n = len(seqs)
C = eng.single_sizes("lz4")
print("packed/bytewise singles", eng.stat("packed_jobs"), eng.stat("bytewise_jobs"))
ref = np.array([olib.ref_lz4f_size(s) for s in seqs])
print("single mismatches", int((C != ref).sum()), "of", n)
S = eng.tile_sizes("lz4", 0, n, 0, n)
print("packed/bytewise pairs", eng.stat("packed_jobs"), eng.stat("bytewise_jobs"))
refS = np.array([[olib.ref_lz4f_size(np.concatenate([a, b])) for b in seqs] for a in seqs])
bad = np.argwhere(S != refS)
print("pair mismatches", len(bad), "of", n * n)
for i, j in bad[:10]: print("  bad", len(seqs[i]), len(seqs[j]), S[i, j], refS[i, j])
eng.set_option("lz4_packed", 0)
S2 = eng.tile_sizes("lz4", 0, n, 0, n)
print("bytewise mismatches", int((S2 != refS).sum()))
eng.set_option("lz4_packed", 1)

# ---- timing: c4-like ----
G = int(os.environ.get("G", "96")); L = int(os.environ.get("L", "5000000"))
import torch
big = [g.cpu().numpy() for g in synth.phylogeny_torch(G, L, 4, torch.device("cuda", 0))]
t = time.time(); eng.upload_sequences(big); print("upload s", time.time() - t)
t = time.time(); Cb = eng.single_sizes("lz4"); dt = time.time() - t
print("singles s", dt, "kernel ms", eng.last_kernel_ms())
print("check single0", Cb[0], olib.ref_lz4f_size(big[0]))
for rows in [G]:
    eng.set_option("invalidate_caches", 1)
    t = time.time(); S = eng.tile_sizes("lz4", 0, rows, 0, G); dt = time.time() - t
    ms = eng.stat("total_kernel_ms"); main = eng.stat("main_kernel_ms")
    lens_b = np.array([len(b) for b in big], dtype=np.float64)
    algo = float(rows * lens_b.sum() + G * lens_b[:rows].sum())
    print(json.dumps({"rows": rows, "cols": G, "jobs": rows * G, "wall_s": dt, "kernel_ms": ms, "main_ms": main,
                      "algo_GBps_main": algo / (main * 1e-3) / 1e9, "pairs_per_s": rows * G / 2 / dt,
                      "packed": eng.stat("packed_jobs")}))
chk = [(0, 1), (3, 2), (G - 1, 0)]
for a, b in chk:
    print("check pair", a, b, S[a, b], olib.ref_lz4f_size(np.concatenate([big[a], big[b]])))

# ---- timing: c3-like ----
N3 = 3000
small = [dna(int(rng.normal(10700, 150)), 900 + i) for i in range(N3)]
eng.upload_sequences(small)
t = time.time(); S3 = eng.tile_sizes("lz4", 0, 768, 0, N3); dt = time.time() - t
print(json.dumps({"c3 jobs": 768 * N3, "wall_s": dt, "kernel_ms": eng.stat("total_kernel_ms"), "main_ms": eng.stat("main_kernel_ms"),
                  "pairs_per_s_main": 768 * N3 / 2 / (eng.stat("main_kernel_ms") * 1e-3), "packed": eng.stat("packed_jobs")}))
for a, b in [(0, 1), (700, 2999), (5, 5)]:
    print("check c3", S3[a, b], olib.ref_lz4f_size(np.concatenate([small[a], small[b]])))

"""Dev script (GPU box): LZ4 parity against the system liblz4 + first timings."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from snacc_b200.engine import Engine
from oracle import lib as olib
from oracle.make_golden import synth_vector, VECTOR_KINDS

eng = Engine(0)
rng = np.random.default_rng(5)
# ---- parity on mixed small corpus ----
seqs = []
for k, kind in enumerate(VECTOR_KINDS):
    for n in [1, 5, 12, 13, 300, 11000, 40000, 65535, 65536, 65537, 70000, 140000]:
        seqs.append(synth_vector(kind, n, 17 * k + n))
eng.upload_sequences(seqs)
t = time.time(); C = eng.single_sizes("lz4"); print("singles ms", eng.last_kernel_ms(), time.time() - t)
ref = np.array([olib.ref_lz4f_size(s) for s in seqs])
print("single mismatches", int((C != ref).sum()), "of", len(seqs))
xs = rng.integers(0, len(seqs), 600); ys = rng.integers(0, len(seqs), 600)
S = eng.pair_sizes("lz4", xs, ys)
refp = np.array([olib.ref_lz4f_size(np.concatenate([seqs[a], seqs[b]])) for a, b in zip(xs, ys)])
print("pair mismatches", int((S != refp).sum()), "of", len(xs))
bad = np.nonzero(S != refp)[0][:10]
for b in bad: print("  bad", xs[b], ys[b], len(seqs[xs[b]]), len(seqs[ys[b]]), S[b], refp[b])

# ---- timing: c4-like (5 Mbp genomes, linked regime) ----
def dna(n, seed): return np.random.default_rng(seed).choice(np.frombuffer(b"ACGT", np.uint8), size=n)
G = 32
L = 5_000_000
big = [dna(L + int(rng.integers(-20000, 20000)), 100 + i) for i in range(G)]
t = time.time(); eng.upload_sequences(big); print("upload s", time.time() - t)
t = time.time(); C = eng.single_sizes("lz4"); print("singles(with prefix) s", time.time() - t, eng.last_kernel_ms())
print("check single0", C[0], olib.ref_lz4f_size(big[0]))
for inflight in [148 * 32, 148 * 64, 148 * 128, 148 * 256]:
    eng.set_option("streams_in_flight", inflight)
    nrow = max(1, inflight // G)
    nrow = min(nrow, G)
    xs = np.repeat(np.arange(nrow), G); ys = np.tile(np.arange(G), nrow)
    # enough jobs to fill `inflight` streams twice
    reps = max(1, (2 * inflight) // xs.size)
    xs = np.tile(xs, reps); ys = np.tile(ys, reps)
    t = time.time(); S = eng.pair_sizes("lz4", xs, ys); dt = time.time() - t
    ms, nl = eng.last_kernel_ms()
    algo_bytes = float(sum(len(big[a]) + len(big[b]) for a, b in zip(xs, ys)))
    print(json.dumps({"inflight": inflight, "jobs": int(xs.size), "kernel_ms": ms, "wall_s": dt,
                      "algo_GBps": algo_bytes / (ms * 1e-3) / 1e9, "pairs_per_s": xs.size / (ms * 1e-3)}))
print("check pair(0,1)", S[1], olib.ref_lz4f_size(np.concatenate([big[0], big[1]])))

# ---- timing: c3-like (11 kbp genomes, single-block regime) ----
N3 = 2000
small = [dna(int(rng.normal(10700, 150)), 900 + i) for i in range(N3)]
eng.upload_sequences(small)
for inflight in [148 * 64, 148 * 256, 148 * 1024]:
    eng.set_option("streams_in_flight", inflight)
    nj = 400_000
    xs = rng.integers(0, N3, nj).astype(np.int32); xs.sort(); ys = rng.integers(0, N3, nj).astype(np.int32)
    t = time.time(); S = eng.pair_sizes("lz4", xs, ys); dt = time.time() - t
    ms, nl = eng.last_kernel_ms()
    algo_bytes = float(sum(len(small[a]) + len(small[b]) for a, b in zip(xs[:1000], ys[:1000]))) / 1000 * nj
    print(json.dumps({"c3 inflight": inflight, "jobs": nj, "kernel_ms": ms, "wall_s": dt,
                      "algo_GBps": algo_bytes / (ms * 1e-3) / 1e9, "pairs_per_s": nj / (ms * 1e-3)}))
k = 12345
print("check c3 pair", S[k], olib.ref_lz4f_size(np.concatenate([small[xs[k]], small[ys[k]]])))

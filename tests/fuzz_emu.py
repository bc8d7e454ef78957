"""Long-running fuzz of the kernel logic emulated on the CPU (tests/host_emu.cu) against the real liblz4 / zlib:
the singles pass in 64-probe batches, pair streams cut into tile segments, and the chunk-parallel deflate parse, on random
four-symbol texts with skewed base frequencies, planted repeats (short / long, near / far, exact / mutated, overlapping) and
runs.  Not collected by pytest; run by hand:  python tests/fuzz_emu.py SEED SECONDS   (round 2: 6 x 1500 s + 4 x 420 s, no
mismatch).  Needs tests/libhost_emu.so (built by tests/test_host_emu.py)."""
import ctypes, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from oracle import lib
e = ctypes.CDLL(os.path.join(HERE, 'libhost_emu.so'))
e.emu_lz4_packed.restype = ctypes.c_int64
e.emu_lz4_packed.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
e.emu_set_batch.argtypes = [ctypes.c_int]; e.emu_set_segments.argtypes = [ctypes.c_int]
e.emu_deflate_chunked.restype = ctypes.c_int64
e.emu_deflate_chunked.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
A = np.frombuffer(b"ACGT", dtype=np.uint8)
def gen(n):
    probs = rng.dirichlet(np.ones(4) * rng.choice([0.3, 1, 5]))
    x = A[rng.choice(4, size=n, p=probs)].copy()
    for _ in range(int(rng.integers(0, 40))):       # planted repeats: short/long, near/far, some mutated
        L = int(rng.choice([5, 8, 12, 13, 20, 64, 300, 3000])); 
        if n <= L + 10: continue
        dst = int(rng.integers(L, n - L)); dist = int(rng.choice([1, 2, 3, 4, 7, 100, 5000, 65530, 65536, 70000]))
        src = dst - dist
        if src < 0: continue
        for i in range(L): x[dst + i] = x[src + i]      # overlapping copy semantics
        if rng.random() < 0.5: x[dst + L // 2] = A[int(rng.integers(4))]
    if rng.random() < 0.2:
        a = int(rng.integers(0, n - 1)); x[a:a + int(rng.integers(1, 5000))] = A[int(rng.integers(4))]
    return x
def call(x, y=None):
    return e.emu_lz4_packed(x.ctypes.data, len(x), y.ctypes.data if y is not None else None, len(y) if y is not None else -1)
t0 = time.time(); bad = 0; n_l = n_s = n_d = 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
while time.time() - t0 < budget:
    # singles through the batches
    n = int(rng.choice([70000, 131072, 140000, 200000, 65537, 66000, 400000]))
    x = gen(n)
    e.emu_set_batch(1)
    if call(x) != lib.ref_lz4f_size(x): bad += 1; print("BATCH MISMATCH", n, flush=True); np.save(f"/tmp/bad_batch_{n_l}.npy", x)
    e.emu_set_batch(0); n_l += 1
    # pair streams in segments
    y = gen(int(rng.choice([200000, 300000, 500000])))
    k = int(rng.choice([2, 3, 4, 8])); e.emu_set_segments(k)
    xs = x[:int(rng.choice([100, 70000, len(x)]))]
    if call(xs, y) != lib.ref_lz4f_size(np.concatenate([xs, y])): bad += 1; print("SEGMENT MISMATCH", k, len(xs), len(y), flush=True)
    e.emu_set_segments(1); n_s += 1
    # chunked deflate
    z = gen(int(rng.integers(131072, 180000)))
    info = (ctypes.c_int32 * 2)()
    for level in (9, 6):
        got = e.emu_deflate_chunked(z.ctypes.data, len(z), level, info)
        if got != lib.ref_deflate_size(z, level): bad += 1; print("DEFLATE MISMATCH", level, len(z), got, list(info), flush=True); np.save(f"/tmp/bad_dfl_{n_d}.npy", z)
        n_d += 1
print("cases", n_l, n_s, n_d, "bad", bad, "time", round(time.time() - t0))

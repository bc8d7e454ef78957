"""CPU: the product's __host__ __device__ parse functions (snacc_b200/csrc/*.cuh), compiled for the host
by tests/host_emu.cu, against the oracle.  This checks the exact kernel logic -- two-segment stream
accessor, prefix checkpoint, block budget rules -- without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import lib
from oracle.make_golden import VECTOR_KINDS, synth_vector
from snacc_b200 import _build

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "libhost_emu.so")
    src = os.path.join(HERE, "host_emu.cu")
    deps = [src] + _build.HEADERS
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([_build._nvcc(), "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                               "-shared", "-o", so, src])
    e = ctypes.CDLL(so)
    e.emu_lz4_size.restype = ctypes.c_int64
    e.emu_lz4_size.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
    e.emu_lz4_packed.restype = ctypes.c_int64
    e.emu_lz4_packed.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
    e.emu_lz4_exact.restype = ctypes.c_int64
    e.emu_lz4_exact.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
    e.emu_lz4_packed_exc.restype = ctypes.c_int64
    e.emu_lz4_packed_exc.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64]
    e.emu_deflate_size.restype = ctypes.c_int64
    e.emu_deflate_size.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    e.emu_flush_compact_fuzz.restype = ctypes.c_int64
    e.emu_flush_compact_fuzz.argtypes = [ctypes.c_uint32, ctypes.c_int32, ctypes.c_void_p]
    e.emu_deflate_size_ex.restype = ctypes.c_int64
    e.emu_deflate_size_ex.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p]
    e.emu_deflate_chunked.restype = ctypes.c_int64
    e.emu_deflate_chunked.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
    return e


def lz4_emu(emu, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.uint8)
    if y is None:
        return emu.emu_lz4_size(x.ctypes.data, x.size, None, -1)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    return emu.emu_lz4_size(x.ctypes.data, x.size, y.ctypes.data, y.size)


@pytest.mark.parametrize("kind", VECTOR_KINDS)
def test_lz4_kernel_logic_singles_and_pairs(emu, kind):
    rng = np.random.default_rng(11)
    for lx in [1, 12, 13, 700, 11000, 65535, 65536, 65537, 70000, 131072, 150000]:
        x = synth_vector(kind, lx, lx + 1)
        assert lz4_emu(emu, x) == lib.lz4f_size(x), (kind, lx)
        for ly in [1, 13, 9000, 65536, 80000]:
            y = synth_vector(kind if rng.random() < 0.7 else "dna", ly, ly + 7)
            assert lz4_emu(emu, x, y) == lib.lz4f_size(np.concatenate([x, y])), (kind, lx, ly)


def test_lz4_related_genomes_cross_boundary_matches(emu):
    from snacc_b200 import synth
    g = synth.phylogeny(4, 40000, seed=3)
    for a in g:
        for b in g:
            assert lz4_emu(emu, a, b) == lib.ref_lz4f_size(np.concatenate([a, b]))


# ---- packed (2-bit) LZ4 tile-kernel logic: ring refills, global fallback, prefix checkpoints, turbo loop ----
def _call2(fn, x, y, *extra):
    x = np.ascontiguousarray(x, dtype=np.uint8)
    if y is None:
        return fn(x.ctypes.data, x.size, None, -1, *extra)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    return fn(x.ctypes.data, x.size, y.ctypes.data, y.size, *extra)


def _dna(n, seed, alphabet=b"ACGT", p=None):
    return np.frombuffer(alphabet, dtype=np.uint8)[np.random.default_rng(seed).choice(len(alphabet), size=n, p=p)]


def _four_symbol_vector(kind, n, seed):
    if kind == "run":
        return np.full(n, 65, np.uint8)
    if kind == "period":
        return np.frombuffer((b"ACGTTGCA" * (n // 8 + 1))[:n], dtype=np.uint8).copy()
    if kind == "two":
        return _dna(n, seed, b"AT")
    if kind == "skew":
        return _dna(n, seed, b"ACGT", [0.85, 0.05, 0.05, 0.05])
    if kind == "lower":
        return _dna(n, seed, b"acgt")
    if kind == "repeat":
        return np.tile(_dna(max(1, n // 7), seed), 8)[:n].copy()
    if kind == "longrep":
        u = _dna(max(1, n // 2 + 3), seed)
        return np.concatenate([u, u])[:n].copy()
    return _dna(n, seed)


def test_lz4_packed_logic_regime_boundaries(emu):
    bad = []
    for lx in [1, 5, 12, 13, 20, 700, 11000, 40000, 65519, 65520, 65535, 65536, 65537, 65540, 70000, 131071, 131072, 131073,
               150000, 300000]:
        x = _dna(lx, lx)
        if _call2(emu.emu_lz4_packed, x, None) != lib.lz4f_size(x):
            bad.append(("single", lx))
        for ly in [16, 17, 100, 9000, 25000, 65536, 80000, 200000]:
            y = _dna(ly, ly + 3)
            if _call2(emu.emu_lz4_packed, x, y) != lib.ref_lz4f_size(np.concatenate([x, y])):
                bad.append(("pair", lx, ly))
    assert not bad, bad[:10]


@pytest.mark.parametrize("kind", ["run", "period", "two", "skew", "lower", "repeat", "longrep"])
def test_lz4_packed_logic_adversarial_four_symbol_inputs(emu, kind):
    bad = []
    for lx in [1, 13, 300, 11000, 65535, 65536, 65537, 100000, 140000]:
        x = _four_symbol_vector(kind, lx, lx)
        if _call2(emu.emu_lz4_packed, x, None) != lib.ref_lz4f_size(x):
            bad.append(("single", lx))
        for ly in [16, 300, 30000, 65536, 100000]:
            for k2 in (kind, "skew"):
                y = _four_symbol_vector(k2, ly, ly + 5)
                z = np.concatenate([x, y])
                if len(set(z.tolist())) > 4:
                    continue
                if _call2(emu.emu_lz4_packed, x, y) != lib.ref_lz4f_size(z):
                    bad.append(("pair", k2, lx, ly))
        got = _call2(emu.emu_lz4_packed, x, x)
        if got != lib.ref_lz4f_size(np.concatenate([x, x])) and not (lx < 16 and got == -1):
            bad.append(("self", lx))
    assert not bad, bad[:10]


def test_lz4_packed_related_genomes(emu):
    from snacc_b200 import synth
    g = synth.phylogeny(4, 40000, seed=3) + synth.phylogeny(3, 200000, seed=4)
    for a in g:
        for b in g:
            assert _call2(emu.emu_lz4_packed, a, b) == lib.ref_lz4f_size(np.concatenate([a, b]))


def test_lz4_packed_tile_segments(emu):
    """PkSegStore / PkRing::restart: a pair stream cut at ring-refill boundaries and continued from its saved table,
    epoch plane and state on a rebuilt ring gives liblz4's size (what lz4_pk_pair_kernel does with n_seg > 1)"""
    from snacc_b200 import synth
    emu.emu_set_segments.argtypes = [ctypes.c_int]
    g = synth.phylogeny(3, 420000, seed=11)
    cases = [(g[0], g[1]), (g[2][:70000], g[0]), (g[1][:100], np.concatenate([g[2], g[0]])),
             (_four_symbol_vector("longrep", 150000, 5), _four_symbol_vector("repeat", 300000, 6))]
    try:
        for k in (2, 3, 4, 8):
            emu.emu_set_segments(k)
            for x, y in cases:
                assert _call2(emu.emu_lz4_packed, x, y) == lib.ref_lz4f_size(np.concatenate([x, y])), (k, len(x), len(y))
    finally:
        emu.emu_set_segments(1)


def test_lz4_singles_batches_of_32_probes(emu):
    """pk_batch_run (the singles kernel's warp-wide batches: 32 probes looked up at once, the chain walked over their
    results, stale lanes detected) against liblz4: singles in both regimes and the single-block pair streams, random,
    repetitive and related inputs; and it is the path that does the work"""
    from snacc_b200 import synth
    emu.emu_set_batch.argtypes = [ctypes.c_int]
    emu.emu_lz4_step_counts.argtypes = [ctypes.c_void_p]
    cnt0, cnt1 = (ctypes.c_uint64 * 3)(), (ctypes.c_uint64 * 3)()
    bad = []
    emu.emu_set_batch(1)
    try:
        for lx in [1, 13, 40, 700, 11000, 65535, 65536, 65537, 70000, 131073, 300000]:
            for kind in ["rand", "run", "period", "two", "skew", "repeat", "longrep"]:
                x = _dna(lx, lx) if kind == "rand" else _four_symbol_vector(kind, lx, lx)
                if _call2(emu.emu_lz4_packed, x, None) != lib.ref_lz4f_size(x):
                    bad.append(("single", kind, lx))
            for ly in [16, 300, 9000, 30000]:
                x, y = _dna(lx, lx), _dna(ly, ly + 3)
                if lx + ly <= 65536 and _call2(emu.emu_lz4_packed, x, y) != lib.ref_lz4f_size(np.concatenate([x, y])):
                    bad.append(("pair", lx, ly))
        g = synth.phylogeny(3, 500000, seed=8)
        emu.emu_lz4_step_counts(cnt0)
        for a in g:
            if _call2(emu.emu_lz4_packed, a, None) != lib.ref_lz4f_size(a):
                bad.append(("genome", len(a)))
        emu.emu_lz4_step_counts(cnt1)
        for a in g[:2]:
            for b in g[:2]:
                if _call2(emu.emu_lz4_packed, a, b) != lib.ref_lz4f_size(np.concatenate([a, b])):
                    bad.append(("genome pair",))
    finally:
        emu.emu_set_batch(0)
    assert not bad, bad[:10]
    general, scalar, batch = (int(cnt1[i] - cnt0[i]) for i in range(3))
    assert batch > 50 * (general + scalar), (general, scalar, batch)      # singles of genomes: ~99.9 % of the probes


@pytest.mark.parametrize("kind", VECTOR_KINDS)
def test_lz4_byte_exact_step_on_any_bytes(emu, kind):
    """pk_step_exact alone (the path probes near non-alphabet bytes take): true-byte hash, alphabet buckets in the slot
    table, the rest in the overflow table, DETECT checkpoint + resume on the 17-bit table -- against liblz4"""
    rng = np.random.default_rng(7)
    bad = []
    for lx in [13, 700, 40000, 65536, 65537, 70000, 140000]:
        x = synth_vector(kind, lx, lx + 1)
        if lx > 65536 and _call2(emu.emu_lz4_exact, x, None) != lib.ref_lz4f_size(x):
            bad.append(("single", lx))
        for ly in [16, 17, 9000, 65536, 80000, 150000]:       # (y < 16: host-side eligibility rule of the packed pair path)
            if lx + ly <= 65536:
                continue
            y = synth_vector(kind if rng.random() < 0.6 else "dna", ly, ly + 7)
            got = _call2(emu.emu_lz4_exact, x, y)
            if got != lib.ref_lz4f_size(np.concatenate([x, y])) and got != -1:
                bad.append(("pair", lx, ly, got))
    assert not bad, bad[:10]


def _sprinkle(seq, rate, seed, alt=b"NNNNRYKMSWacgtn"):
    """genome-like input: a fraction `rate` of the bases replaced by N / IUPAC / lower-case bytes, plus a few runs of N"""
    rng = np.random.default_rng(seed)
    s = np.array(seq, dtype=np.uint8, copy=True)
    m = rng.random(s.size) < rate
    s[m] = np.frombuffer(alt, dtype=np.uint8)[rng.integers(0, len(alt), int(m.sum()))]
    for _ in range(int(rng.integers(0, 4))):
        a = int(rng.integers(0, max(1, s.size - 1)))
        s[a:a + int(rng.integers(1, 700))] = ord("N")
    return s


@pytest.mark.parametrize("rate", [1e-4, 1e-3, 2e-2])
def test_lz4_packed_path_with_flagged_bases(emu, rate):
    """the EXC kernels' logic: fast loop up to the flagged granules, byte-exact crossings, forced mismatches inside
    candidate windows, overflow table, N runs, exceptions at the x|y junction and at block / ring-refill boundaries"""
    from snacc_b200 import synth
    bad = []
    g = synth.phylogeny(3, 180000, seed=21) + [_dna(70000, 5), _dna(300000, 6)]
    seqs = [_sprinkle(s, rate, 30 + i) for i, s in enumerate(g)]
    seqs[1][-3:] = ord("N"); seqs[2][:2] = ord("n"); seqs[3][65534:65540] = ord("N"); seqs[4][131070:131075] = ord("R")
    for i, x in enumerate(seqs):
        if x.size > 65536 and _call2(emu.emu_lz4_packed_exc, x, None) != lib.ref_lz4f_size(x):
            bad.append(("single", i))
        for j, y in enumerate(seqs):
            got = _call2(emu.emu_lz4_packed_exc, x, y)
            if got != lib.ref_lz4f_size(np.concatenate([x, y])):
                bad.append(("pair", i, j, got))
    # clean inputs take the same path unchanged
    assert _call2(emu.emu_lz4_packed_exc, g[0], g[1]) == lib.ref_lz4f_size(np.concatenate([g[0], g[1]]))
    assert not bad, bad[:10]


def test_lz4_tile_segments_with_flagged_bases(emu):
    """tile segments on the EXC path: the ring AND the dirty ring inside it are rebuilt at a cut (PkRing::restart with the
    EXC cover), table / state / overflow table carry over"""
    from snacc_b200 import synth
    emu.emu_set_segments.argtypes = [ctypes.c_int]
    g = synth.phylogeny(3, 420000, seed=23)
    seqs = [_sprinkle(s, r, 50 + i) for i, (s, r) in enumerate(zip(g, [1e-4, 1e-3, 1e-5]))]
    seqs[0][131070:131075] = ord("N"); seqs[1][-2:] = ord("n")
    try:
        for k in (2, 3, 8):
            emu.emu_set_segments(k)
            for i, j in [(0, 1), (1, 0), (2, 1), (1, 2)]:
                x, y = seqs[i][:150000], seqs[j]
                assert _call2(emu.emu_lz4_packed_exc, x, y) == lib.ref_lz4f_size(np.concatenate([x, y])), (k, i, j)
    finally:
        emu.emu_set_segments(1)


def test_lz4_packed_refuses_more_than_four_symbols(emu):
    assert _call2(emu.emu_lz4_packed, np.frombuffer(b"ACGTNACGTNACGTNACGT", dtype=np.uint8), None) == -2


# ---- deflate: index + F tables + junction + x checkpoint + lazy parse + block cost, as the kernels do it ----
@pytest.mark.parametrize("level", [9, 6])
def test_deflate_logic_boundaries(emu, level):
    bad = []
    for lx in [1, 2, 3, 5, 100, 300, 5000, 32506, 32768, 65274, 65275, 65400, 70000]:
        x = _dna(lx, lx)
        if _call2(emu.emu_deflate_size, x, None, level) != lib.ref_deflate_size(x, level):
            bad.append(("single", lx))
        for ly in [1, 2, 3, 100, 5000, 33000]:
            y = _dna(ly, ly + 1)
            if _call2(emu.emu_deflate_size, x, y, level) != lib.ref_deflate_size(np.concatenate([x, y]), level):
                bad.append(("pair", lx, ly))
    assert not bad, bad[:10]


@pytest.mark.parametrize("kind", VECTOR_KINDS)
def test_deflate_logic_alphabets(emu, kind):
    bad = []
    for level in (9, 6):
        for lx in [13, 300, 9000, 33000]:
            x = synth_vector(kind, lx, lx + 1)
            if _call2(emu.emu_deflate_size, x, None, level) != lib.ref_deflate_size(x, level):
                bad.append(("single", level, lx))
            for ly in [40, 9000, 34000]:
                y = synth_vector(kind, ly, ly + 7)
                if _call2(emu.emu_deflate_size, x, y, level) != lib.ref_deflate_size(np.concatenate([x, y]), level):
                    bad.append(("pair", level, lx, ly))
    assert not bad, bad[:10]


def test_deflate_logic_related_genomes_long_matches_across_the_boundary(emu):
    from snacc_b200 import synth
    g = synth.phylogeny(3, 40000, seed=3)
    for level in (9, 6):
        for a in g:
            for b in g:
                assert _call2(emu.emu_deflate_size, a, b, level) == lib.ref_deflate_size(np.concatenate([a, b]), level)


@pytest.mark.parametrize("seed", range(4))
def test_lz4_packed_17bit_slots_never_resurrect_stale_candidates(emu, seed):
    """the 16-bit + epoch-bit table (PkTab KIND 2) relies on the rolling sweep; without it this input gives wrong sizes"""
    from snacc_b200 import synth
    x, y = synth.stale_slot_stream(seed)
    assert _call2(emu.emu_lz4_packed, x, y) == lib.ref_lz4f_size(np.concatenate([x, y]))


def _deflate_ex(emu, x, y, level, canon):
    info = (ctypes.c_int32 * 4)()
    x = np.ascontiguousarray(x, dtype=np.uint8)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    r = emu.emu_deflate_size_ex(x.ctypes.data, x.size, y.ctypes.data, y.size, level, canon, info)
    return r, info[0], info[1] + 1000 * info[2]      # info[2]: junction words where the continued walk differs from the general one


@pytest.mark.parametrize("level", [9, 6])
def test_deflate_canonical_symbol_stream_shortcut(emu, level):
    """pair streams whose y is long enough take their blocks from the recorded parse of y alone (dfl_pair_stream): the
    synchronisation point, block boundaries that fall inside the shortcut (y > 150 kbp: several 16383-symbol blocks), the
    tail handed back to the serial parser, window-slide phases of x, and inputs for which there is no canonical stream"""
    from snacc_b200 import synth
    cases = []
    for lx, ly in [(5000, 40000), (70000, 37000), (300, 200000), (100000, 400000), (65275, 36864), (65274, 36865),
                   (50000, 36863), (1, 50000), (32506, 45000)]:
        cases.append((_dna(lx, lx), _dna(ly, ly + 1), ly >= 36864))
    rng = np.random.default_rng(9)
    cases.append((rng.integers(65, 67, 30000).astype(np.uint8), rng.integers(65, 67, 120000).astype(np.uint8), True))
    cases.append((rng.integers(65, 81, 30000).astype(np.uint8), rng.integers(65, 81, 90000).astype(np.uint8), None))
    cases.append((rng.integers(0, 256, 30000).astype(np.uint8), rng.integers(0, 256, 90000).astype(np.uint8), False))
    g = synth.phylogeny(2, 120000, seed=3)
    cases += [(g[0], g[1], True), (g[1], g[1], True)]
    bad = []
    for x, y, expect_used in cases:
        ref = lib.ref_deflate_size(np.concatenate([x, y]), level)
        r, used, fell_back = _deflate_ex(emu, x, y, level, 1)
        r0, used0, _ = _deflate_ex(emu, x, y, level, 0)
        if r != ref or r0 != ref or used0 or fell_back or (expect_used is not None and bool(used) != expect_used):
            bad.append((x.size, y.size, ref, r, r0, used, fell_back))
    assert not bad, bad


@pytest.mark.parametrize("seed", range(3))
def test_deflate_block_cost_on_compacted_trees_equals_zlib_tree_construction(emu, seed):
    """dfl_flush_block_compact (Huffman trees built over the nonzero symbols only, the version the pair kernel keeps in
    shared memory) against dfl_flush_block (trees.c restated, itself pinned against zlib through the size tests) on
    random histograms: ties, power-of-two frequencies (length overflow), DNA-like shapes, stored candidates"""
    applied = ctypes.c_int32()
    assert emu.emu_flush_compact_fuzz(seed, 60000, ctypes.byref(applied)) == 0
    assert applied.value > 20000


def test_deflate_shortcuts_on_repetitive_inputs(emu):
    """long single-base runs (matches of 258: the nice_length stop), tandem repeats, a 3 kbp unit repeated (every walk ends
    at nice_length), N runs, x == y: the junction / match-table / canonical-stream shortcuts against zlib, and every
    shortcut word against the general chain walk (second return value)"""
    rng = np.random.default_rng(3)

    def dna(n):
        return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]

    unit, u2 = dna(37), dna(3000)
    rep = np.concatenate([np.tile(u2, 15), dna(40000)])
    cases = [
        (np.concatenate([np.full(40000, 65, np.uint8), dna(20000)]),
         np.concatenate([dna(5000), np.full(60000, 65, np.uint8), dna(30000)])),
        (np.tile(unit, 2000), np.tile(unit, 2500)),
        (np.tile(u2, 20), rep),
        (np.concatenate([dna(30000), np.full(5000, 78, np.uint8), dna(30000)]),
         np.concatenate([np.full(3000, 78, np.uint8), dna(60000), np.full(200, 78, np.uint8), dna(20000)])),
        (rep, rep),
    ]
    bad = []
    for x, y in cases:
        for level in (9, 6):
            ref = lib.ref_deflate_size(np.concatenate([x, y]), level)
            r, _, mism = _deflate_ex(emu, x, y, level, 1)
            if r != ref or mism:
                bad.append((x.size, y.size, level, ref, r, mism))
    assert not bad, bad


@pytest.mark.parametrize("level", [9, 6])
def test_deflate_sequence_alone_in_parallel_chunks(emu, level):
    """dfl_chunk_*_kernel + dfl_alone_kernel emulated: chunks of 8 KiB parsed independently from a 'just emitted a match'
    state fall in with the true parse inside the overlap; the stitched symbol stream equals the serial parse's symbol
    for symbol and the size (blocks from the cumulative rows + serial tail) equals zlib's.  Inputs where neighbouring
    chunks never meet (periodic runs) are given up -- never a wrong stream."""
    from snacc_b200 import synth
    g = synth.phylogeny(2, 300000, seed=31)
    cases = [("genome", g[0], True), ("genome", g[1], True), ("random", _dna(140000, 3), True),
             ("lower", _four_symbol_vector("lower", 200000, 5), True), ("skew", _four_symbol_vector("skew", 160000, 6), None),
             ("repeat", _four_symbol_vector("repeat", 180000, 7), None), ("run", _four_symbol_vector("run", 150000, 8), None),
             ("period", _four_symbol_vector("period", 150000, 9), None),
             ("bytes", np.frombuffer(np.random.default_rng(4).bytes(150000), dtype=np.uint8).copy(), None)]
    info = (ctypes.c_int32 * 2)()
    for name, x, must_stand in cases:
        got = emu.emu_deflate_chunked(x.ctypes.data, len(x), level, info)
        assert got == lib.ref_deflate_size(x, level), (name, len(x), got, list(info))
        if must_stand:
            assert info[0] == 1, (name, list(info))
    short = _dna(1000, 1)
    assert emu.emu_deflate_chunked(short.ctypes.data, len(short), level, info) == -3


@pytest.mark.parametrize("level", [9, 6])
def test_deflate_parallel_chunks_length_boundaries(emu, level):
    """the chunk grid at its edges: the shortest sequence the path takes (16 chunks), lengths where n - 1 KiB falls on, just
    before and just after a chunk boundary (the last chunk is then one to two chunks long), and a repeat that straddles a
    chunk start (a match that begins in one chunk and ends in the next)"""
    info = (ctypes.c_int32 * 2)()
    base = _dna(160000, 77)
    for n in [131072, 131073, 16 * 8192 + 1024 - 1, 16 * 8192 + 1024, 16 * 8192 + 1024 + 1, 17 * 8192 + 1023, 18 * 8192 + 1025]:
        x = base[:n].copy()
        got = emu.emu_deflate_chunked(x.ctypes.data, len(x), level, info)
        assert got == lib.ref_deflate_size(x, level), (n, got, list(info))
        assert info[0] == 1 and info[1] == (n - 1024) // 8192, (n, list(info))
    x = base[:150000].copy()
    for s in (8192 * 3, 8192 * 7, 8192 * 11):                 # 300-byte copies of earlier text across three chunk starts
        x[s - 150:s + 150] = x[s - 5150:s - 4850]
    got = emu.emu_deflate_chunked(x.ctypes.data, len(x), level, info)
    assert got == lib.ref_deflate_size(x, level) and info[0] == 1, (got, list(info))

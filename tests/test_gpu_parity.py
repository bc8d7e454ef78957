"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes Engine), against the oracle and the
golden fixtures written by the unmodified reference.  Integer sizes must be identical; NCD bit-exact."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

from oracle import lib as olib
from oracle import snacc_oracle
from oracle.fasta_shim import COMPLEMENT_TABLE
from oracle.make_golden import VECTOR_KINDS, synth_vector

pytestmark = pytest.mark.gpu

ALGOS = ["lz4", "gzip", "zlib"]


def _ref_len(data, algo):
    return olib.ref_compressed_len(data, algo)


@pytest.fixture(scope="module")
def ref_sizes(golden_dir):
    return json.load(open(os.path.join(golden_dir, "reference_sizes.json")))


@pytest.mark.parametrize("case", ["lz4", "lz4_rc", "gzip", "gzip_rc"])
def test_fixture_matrix_equals_reference_golden(engine, golden_dir, ref_sizes, case):
    """all singles + all ordered pairs of the committed FASTA fixtures == what the reference's own
    compressed_size returned (incl. multi-record reverse complement done on the device)"""
    from snacc_b200.pairwise_ncd import ncd_matrix
    algo, rc = case.split("_")[0], case.endswith("_rc")
    files = [Path(golden_dir) / "fasta" / f for f in ref_sizes["files"]]
    labels, C, S, D = ncd_matrix(files, algo, reverse_complement=rc, engine=engine)
    want = ref_sizes["cases"][case]
    assert (C + 33).tolist() == want["C"]
    assert (S + 33).tolist() == want["S"]
    assert np.array_equal(D, np.array(want["D"]))
    assert np.array_equal(engine.ncd(C, S), np.array(want["D"]))


@pytest.mark.parametrize("algo", ["lz4", "gzip"])
def test_fast_mode_computes_the_upper_triangle_only(engine, golden_dir, ref_sizes, algo):
    """--fast-mode True (README flag; SURVEY.md 8a divergence ledger): S(i, j) for i <= j only, mirrored, NCD from that one
    order -- the sizes it does compute are the reference's"""
    from snacc_b200.pairwise_ncd import ncd_matrix
    from snacc_b200.sharding import ncd_host
    files = [Path(golden_dir) / "fasta" / f for f in ref_sizes["files"]]
    n = len(files)
    labels, C, S, D = ncd_matrix(files, algo, fast_mode=True, engine=engine)
    want = ref_sizes["cases"][algo]
    Sref = np.array(want["S"]) - 33
    iu = np.triu_indices(n, 1)
    Sref[(iu[1], iu[0])] = Sref[iu]
    assert (C + 33).tolist() == want["C"]
    assert np.array_equal(S, Sref)
    assert np.array_equal(D, ncd_host(C, Sref, fast_mode=True))


@pytest.mark.parametrize("case", ["lz4", "gzip_rc"])
def test_cli_csv_is_byte_identical_to_reference_cli(golden_dir, tmp_path, case, monkeypatch):
    from click.testing import CliRunner
    from snacc_b200 import cli as gcli
    algo, rc = case.split("_")[0], case.endswith("_rc")
    d = Path(golden_dir) / "fasta"
    files = [str(f) for f in sorted(d.iterdir()) if not f.name.startswith("big")]
    out = tmp_path / "dist.csv"
    monkeypatch.chdir(tmp_path)
    args = files + ["-o", str(out), "-c", algo, "--no-show-progress"] + (["--reverse-compliment", "True"] if rc else [])
    r = CliRunner().invoke(gcli.cli, args)
    assert r.exit_code == 0, r.output
    got = out.read_text().replace(str(d.absolute()) + "/", "")
    assert got == (Path(golden_dir) / f"reference_cli_{case}.csv").read_text()
    assert (tmp_path / "dist.md").exists()


def test_single_job_shim_matches_reference_signature(golden_dir, ref_sizes):
    import snacc_b200
    f = Path(golden_dir) / "fasta" / "g1.fasta"
    g = Path(golden_dir) / "fasta" / "multi.fa"
    i, j = ref_sizes["files"].index("g1.fasta"), ref_sizes["files"].index("multi.fa")
    assert snacc_b200.compressed_size(f, "lz4") == (f, ref_sizes["cases"]["lz4"]["C"][i])
    assert snacc_b200.compressed_size((f, g), "lz4", reverse_complement=True) == ((f, g), ref_sizes["cases"]["lz4_rc"]["S"][i][j])


@pytest.mark.parametrize("algo", ALGOS)
def test_regime_boundaries_and_alphabets(engine, algo):
    """lengths straddling every regime boundary x adversarial alphabets, singles and random pairs"""
    rng = np.random.default_rng(5)
    sizes = [1, 2, 3, 5, 12, 13, 300, 11000, 40000, 65274, 65535, 65536, 65537, 70000, 131072, 140000]
    seqs = [synth_vector(kind, n, 17 * k + n) for k, kind in enumerate(VECTOR_KINDS) for n in sizes
            if not (algo == "gzip" and n > 70000 and kind in ("run", "period", "low", "nrun"))]
    engine.upload_sequences(seqs)
    C = engine.single_sizes(algo)
    ref = np.array([_ref_len(s, algo) for s in seqs])
    assert np.array_equal(C, ref), np.nonzero(C != ref)[0][:10]
    m = 400 if algo == "lz4" else 150
    xs, ys = rng.integers(0, len(seqs), m), rng.integers(0, len(seqs), m)
    S = engine.pair_sizes(algo, xs, ys)
    refp = np.array([_ref_len(np.concatenate([seqs[a], seqs[b]]), algo) for a, b in zip(xs, ys)])
    bad = np.nonzero(S != refp)[0]
    assert bad.size == 0, [(int(xs[b]), int(ys[b]), seqs[xs[b]].size, seqs[ys[b]].size, int(S[b]), int(refp[b])) for b in bad[:5]]


def test_on_device_reverse_complement(engine):
    rng = np.random.default_rng(2)
    alpha = np.frombuffer(b"ACGTacgtNnRYKMSWBDHVUu-*", dtype=np.uint8)
    seqs, recs = [], []
    for i in range(5):
        rl = [int(v) for v in rng.integers(0, 400, size=int(rng.integers(1, 5)))]
        rl[0] += 1
        seqs.append(rng.choice(alpha, size=sum(rl)))
        recs.append(rl)
    engine.upload_sequences(seqs, reverse_complement=True, records=recs)
    for i, (s, rl) in enumerate(zip(seqs, recs)):
        raw, parts, pos = s.tobytes(), [], 0
        for r in rl:
            parts.append(raw[pos:pos + r].translate(COMPLEMENT_TABLE)[::-1])
            pos += r
        assert engine.download_sequence(i).tobytes() == b"".join(parts)


def test_empty_sequence_is_value_error(engine):
    with pytest.raises(ValueError):
        engine.upload_sequences([b"ACGT", b""])


def test_unsupported_codec_is_key_error(engine):
    engine.upload_sequences([b"ACGTACGTACGT"])
    with pytest.raises(KeyError):
        engine.single_sizes("lzma")


@pytest.mark.parametrize("algo", ALGOS)
def test_small_stream_config_sample(engine, algo):
    """c3 shape: dengue-sized (~11 kbp) genomes, random sample of pair jobs + all singles"""
    from snacc_b200 import synth
    g = synth.phylogeny(64, 10700, seed=3)
    engine.upload_sequences(g)
    C = engine.single_sizes(algo)
    assert np.array_equal(C, np.array([_ref_len(s, algo) for s in g]))
    rng = np.random.default_rng(3)
    m = 3000 if algo == "lz4" else 300
    xs, ys = rng.integers(0, 64, m), rng.integers(0, 64, m)
    S = engine.pair_sizes(algo, xs, ys)
    chk = rng.choice(m, size=min(m, 300), replace=False)
    for k in chk:
        assert S[k] == _ref_len(np.concatenate([g[xs[k]], g[ys[k]]]), algo)


def test_full_size_genomes_lz4(engine):
    """c4 shape at full length (5 Mbp): tile of ordered pairs vs the real liblz4, plus the size-independent
    properties: prefix-checkpoint path == from-scratch path (single of a concatenated upload) and
    determinism under a different number of streams in flight."""
    from snacc_b200 import synth
    g = synth.phylogeny(6, 5_000_000, seed=4, n_indels=2)
    engine.upload_sequences(g)
    C = engine.single_sizes("lz4")
    assert np.array_equal(C, np.array([olib.ref_lz4f_size(s) for s in g]))
    S = engine.tile_sizes("lz4", 0, 3, 0, 6)
    ref = np.array([[olib.ref_lz4f_size(np.concatenate([g[i], g[j]])) for j in range(6)] for i in range(3)])
    assert np.array_equal(S, ref)
    engine.set_option("streams_in_flight", 8)
    engine.set_option("invalidate_caches", 1)
    assert np.array_equal(engine.tile_sizes("lz4", 0, 3, 0, 6), ref)
    engine.set_option("streams_in_flight", 0)
    # concatenation uploaded as ONE sequence takes the no-checkpoint route through the same kernel
    engine.upload_sequences([np.concatenate([g[0], g[1]])])
    assert engine.single_sizes("lz4")[0] == ref[0, 1]
    D = engine.ncd(C[:3], S[:, :3])
    assert np.array_equal(D, snacc_oracle.ncd_from_sizes(C[:3], S[:, :3]))


def test_packed_and_bytewise_lz4_paths_agree(engine):
    """the 2-bit tile kernels and the byte-wise kernels are two independent implementations of the same
    frame size; mixed corpus (a sequence with N falls back to the byte-wise path inside the same call)"""
    from snacc_b200 import synth
    g = synth.phylogeny(12, 150000, seed=9) + [synth_vector("nrun", 90000, 5)]
    engine.upload_sequences(g)
    n = len(g)
    S = engine.tile_sizes("lz4", 0, n, 0, n)
    assert engine.stat("packed_jobs") == 144 and engine.stat("bytewise_jobs") == n * n - 144
    engine.set_option("lz4_packed", 0)
    try:
        S2 = engine.tile_sizes("lz4", 0, n, 0, n)
    finally:
        engine.set_option("lz4_packed", 1)
    assert np.array_equal(S, S2)
    chk = np.random.default_rng(1).integers(0, n, size=(40, 2))
    for i, j in chk:
        assert S[i, j] == olib.ref_lz4f_size(np.concatenate([g[i], g[j]]))


def test_megabase_genomes_gzip(engine):
    """c2 shape (scaled to 1.2 Mbp so that zlib finishes in seconds): all singles, a tile of pairs, warm caches"""
    from snacc_b200 import synth
    g = synth.phylogeny(4, 1_200_000, seed=6, n_indels=2)
    engine.upload_sequences(g)
    C = engine.single_sizes("gzip")
    assert np.array_equal(C, np.array([_ref_len(s, "gzip") for s in g]))
    S = engine.tile_sizes("gzip", 0, 2, 0, 4)
    ref = np.array([[_ref_len(np.concatenate([g[i], g[j]]), "gzip") for j in range(4)] for i in range(2)])
    assert np.array_equal(S, ref)
    assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, 4), ref)          # checkpoints and tables reused
    engine.set_option("invalidate_caches", 1)
    assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, 4), ref)          # and rebuilt
    assert engine.stat("deflate_serial_jobs") == 0                             # every pair stream used the canonical stream of y
    engine.set_option("deflate_canonical", 0)                                  # the full serial parse gives the same sizes
    try:
        assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, 4), ref)
    finally:
        engine.set_option("deflate_canonical", 1)
    engine.set_option("deflate_junction", 2)                                   # general junction walk instead of the smem kernel
    try:
        assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, 4), ref)
    finally:
        engine.set_option("deflate_junction", 3)
    engine.set_option("deflate_index6", 0)                                     # match tables from the 3-byte chain walk only
    engine.set_option("invalidate_caches", 1)
    try:
        assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, 4), ref)
        assert np.array_equal(engine.single_sizes("gzip"), C)
    finally:
        engine.set_option("deflate_index6", 1)
        engine.set_option("invalidate_caches", 1)
    Sz = engine.tile_sizes("zlib", 1, 2, 0, 3)
    refz = np.array([[_ref_len(np.concatenate([g[i], g[j]]), "zlib") for j in range(3)] for i in (1, 2)])
    assert np.array_equal(Sz, refz)
    assert np.array_equal(engine.single_sizes("zlib"), np.array([_ref_len(s, "zlib") for s in g]))


def test_full_size_genomes_gzip(engine):
    """c5 shape: 5 Mbp genomes, gzip level 9 -- every shortcut of the deflate path at full size (≈ 50 blocks per stream,
    ≈ 35 of them taken from the canonical symbol stream, several window slides inside the junction) against zlib 1.3
    run on all host cores; and the same sizes with the shortcuts switched off"""
    from snacc_b200 import synth
    g = synth.phylogeny(3, 5_000_000, seed=5)
    n = len(g)
    engine.upload_sequences(g)
    C = engine.single_sizes("gzip")
    S = engine.tile_sizes("gzip", 0, n, 0, n)
    assert engine.stat("deflate_serial_jobs") == 0
    corpus = np.concatenate(g)
    so = np.zeros(n + 1, dtype=np.uint64)
    so[1:] = np.cumsum([x.size for x in g])
    xs = np.concatenate([np.repeat(np.arange(n, dtype=np.int32), n), np.arange(n, dtype=np.int32)])
    ys = np.concatenate([np.tile(np.arange(n, dtype=np.int32), n), np.full(n, -1, dtype=np.int32)])      # y < 0: x alone
    threads = len(os.sched_getaffinity(0))
    sizes = olib.ref_batch_sizes(corpus, so, xs, ys, "gzip", threads)
    ref = sizes[:n * n].reshape(n, n)
    assert np.array_equal(S, ref)
    assert np.array_equal(C, sizes[n * n:])
    for opt in ("deflate_canonical", "deflate_index6"):
        engine.set_option(opt, 0)
        engine.set_option("invalidate_caches", 1)
        try:
            assert np.array_equal(engine.tile_sizes("gzip", 0, 2, 0, n), ref[:2])
        finally:
            engine.set_option(opt, 1)
    engine.set_option("invalidate_caches", 1)


@pytest.mark.parametrize("algo", ALGOS)
def test_repetitive_inputs(engine, algo):
    """single-base runs of 40-60 k, a 37-base tandem repeat, a 3 kbp unit repeated, N runs: one hash bucket holds almost
    every position (radix sort, head/tail packs, chain limits, nice_length stops, LZ4 matches of tens of kilobases)"""
    rng = np.random.default_rng(3)

    def dna(n):
        return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]

    unit, u2 = dna(37), dna(3000)
    seqs = [np.concatenate([np.full(40000, 65, np.uint8), dna(20000)]),
            np.concatenate([dna(5000), np.full(60000, 65, np.uint8), dna(30000)]),
            np.tile(unit, 2000), np.tile(unit, 2500), np.tile(u2, 20),
            np.concatenate([np.tile(u2, 15), dna(40000)]),
            np.concatenate([np.full(3000, 78, np.uint8), dna(60000), np.full(200, 78, np.uint8), dna(20000)])]
    n = len(seqs)
    engine.upload_sequences(seqs)
    C = engine.single_sizes(algo)
    S = engine.tile_sizes(algo, 0, n, 0, n)
    assert np.array_equal(C, np.array([_ref_len(s, algo) for s in seqs]))
    ref = np.array([[_ref_len(np.concatenate([a, b]), algo) for b in seqs] for a in seqs])
    assert np.array_equal(S, ref)


def test_lz4_stale_table_slots(engine):
    """A/T-only stretches of 66 k - 300 k bases between ACGT stretches: slots of k-mers with C/G age far beyond the
    131072 positions the 17-bit slot encoding can tell apart; the rolling sweep must have retired them"""
    from snacc_b200 import synth
    seqs = []
    for seed in range(3):
        x, y = synth.stale_slot_stream(seed)
        seqs += [x, y]
    engine.upload_sequences(seqs)
    n = len(seqs)
    S = engine.tile_sizes("lz4", 0, n, 0, n)
    assert engine.stat("packed_jobs") == n * n
    ref = np.array([[olib.ref_lz4f_size(np.concatenate([a, b])) for b in seqs] for a in seqs])
    assert np.array_equal(S, ref)

"""CPU: host-side mirror of the reference interface -- FASTA parsing, file discovery, NCD epilogue, CSV,
CLI flag surface, and the world_size-2 row sharding over the gloo backend."""
import os
import socket
from pathlib import Path

import numpy as np
import pytest

from oracle import snacc_oracle
from snacc_b200 import cli as gcli
from snacc_b200 import fasta, pairwise_ncd, sharding


def test_fasta_parser_matches_the_oracle_shim(golden_dir):
    for f in sorted((Path(golden_dir) / "fasta").iterdir()):
        data, recs = fasta.read_fasta(f)
        assert data.tobytes().decode() == snacc_oracle.extract_sequences(f, False), f.name
        assert pairwise_ncd.extract_sequences(f, False) == snacc_oracle.extract_sequences(f, False)
        assert pairwise_ncd.extract_sequences(f, True) == snacc_oracle.extract_sequences(f, True)
        assert sum(recs) == data.size


def test_multi_record_offsets(golden_dir):
    f = Path(golden_dir) / "fasta" / "multi.fa"
    data, so, ro = fasta.load_corpus([f, f])
    assert so.tolist() == [0, 850, 1700] and ro.tolist() == [0, 400, 700, 850, 1250, 1550, 1700]


def test_empty_fasta_raises_value_error(tmp_path):
    p = tmp_path / "empty.fa"
    p.write_text("no header line here\nACGT\n")
    with pytest.raises(ValueError):
        fasta.load_corpus([p])
    with pytest.raises(ValueError):
        pairwise_ncd.extract_sequences(p)


def test_unknown_algorithm_is_a_key_error(golden_dir):
    f = Path(golden_dir) / "fasta" / "tiny.fa"
    with pytest.raises(KeyError):
        pairwise_ncd.compressed_size(f, "snappy")
    with pytest.raises(KeyError):
        pairwise_ncd.compressed_size(f, "lzma")      # valid for the reference, not on the GPU path: no fallback


def test_compute_distance_matches_reference_branches():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        x, y = (int(v) for v in rng.integers(50, 10_000_000, 2))
        cxy, cyx = (int(v) for v in rng.integers(max(x, y), x + y + 100, 2))
        assert pairwise_ncd.compute_distance(x, y, cxy, cyx) == snacc_oracle.compute_distance(x, y, cxy, cyx)
    assert pairwise_ncd.compute_distance(1174721, 1173133, 1242873, 1242873) == 0.05936728806244206


def test_vectorised_epilogue_is_bit_identical_to_the_scalar_formula():
    rng = np.random.default_rng(1)
    n = 23
    C = rng.integers(1000, 5_000_000, n)
    S = rng.integers(1000, 9_000_000, (n, n))
    assert np.array_equal(sharding.ncd_host(C, S), snacc_oracle.ncd_from_sizes(C, S))


def test_file_discovery_and_ordering(golden_dir, tmp_path):
    d = Path(golden_dir) / "fasta"
    (tmp_path / "notes.txt").write_text("x")
    got = gcli.collect_files([str(d), str(d / "g1.fasta")])
    assert got == snacc_oracle.sorted_files([d, d / "g1.fasta"])
    assert [p.name for p in got][:3] == ["big1.fasta", "big2.fasta", "g1.fasta"]


def test_csv_writer_matches_reference_layout(golden_dir, tmp_path):
    d = Path(golden_dir) / "fasta"
    files = [f for f in gcli.collect_files([str(d)]) if not f.name.startswith("big")]
    C, S = snacc_oracle.size_tables(files, "lz4")
    out = tmp_path / "x.csv"
    gcli.write_distance_csv(files, sharding.ncd_host(C, S), out)
    got = out.read_text().replace(str(d.absolute()) + "/", "")
    assert got == (Path(golden_dir) / "reference_cli_lz4.csv").read_text()


def test_cli_flag_surface():
    names = {o for p in gcli.cli.params for o in p.opts}
    for flag in ("-d", "-o", "-n", "-c", "-r", "-f", "-s", "--fast-mode", "--reverse-compliment", "--reverse_complement",
                 "--log", "--show-progress"):
        assert flag in names, flag


def test_cli_rejects_unsupported_codec(golden_dir, tmp_path):
    from click.testing import CliRunner
    r = CliRunner().invoke(gcli.cli, [str(Path(golden_dir) / "fasta" / "tiny.fa"), "-o", str(tmp_path / "o.csv"),
                                      "-c", "lzma"])
    assert r.exit_code != 0 and "not supported on the GPU path" in r.output


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_size_fn(algorithm):
    def fn(data, so, ro, rc, cols):
        from oracle import lib
        from oracle.fasta_shim import COMPLEMENT_TABLE
        raw = data.tobytes()
        seqs = []
        so_l, ro_l = [int(v) for v in so], [int(v) for v in ro]
        for i in range(len(so_l) - 1):
            if rc:
                parts = [raw[a:b].translate(COMPLEMENT_TABLE)[::-1] for a, b in zip(ro_l[:-1], ro_l[1:])
                         if so_l[i] <= a and b <= so_l[i + 1] and b > a]
                seqs.append(b"".join(parts))
            else:
                seqs.append(raw[so_l[i]:so_l[i + 1]])
        C = np.array([lib.compressed_len(s, algorithm) for s in seqs], dtype=np.int64)
        S = np.array([[lib.compressed_len(s + seqs[c], algorithm) for c in cols] for s in seqs],
                     dtype=np.int64).reshape(len(seqs), len(cols))
        return C, S
    return fn


def _worker(rank, world, port, files, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        labels, C, S, D = sharding.all_pairs(files, "lz4", True, False, _size_fn=_oracle_size_fn("lz4"))
        _, Cf, Sf, Df = sharding.all_pairs(files, "lz4", True, True, _size_fn=_oracle_size_fn("lz4"))     # fast mode
        q.put((rank, C.tolist(), S.tolist(), D.tolist(), Sf.tolist(), Df.tolist()))
    finally:
        dist.destroy_process_group()


def test_column_sharding_world_size_2_gloo(golden_dir):
    """N > 1 path on CPU: every rank parses its band of the files, the bands are all-gathered, every rank takes a
    column band (or, in fast mode, its share of the upper-triangle columns), one gather at the end."""
    import torch.multiprocessing as mp
    d = Path(golden_dir) / "fasta"
    files = [f for f in gcli.collect_files([str(d)]) if not f.name.startswith("big")]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C1, S1 = snacc_oracle.size_tables(files, "lz4", True)
    iu = np.triu_indices(len(files), 1)
    S1f = S1.copy()
    S1f[(iu[1], iu[0])] = S1f[iu]
    for rank, C, S, D, Sf, Df in res:
        assert np.array_equal(np.array(C), C1) and np.array_equal(np.array(S), S1)
        assert np.array_equal(np.array(D), snacc_oracle.ncd_from_sizes(C1, S1))
        assert np.array_equal(np.array(Sf), S1f) and np.array_equal(np.array(Df), sharding.ncd_host(C1, S1f, True))


def test_owned_cols_partition():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 512):
        for w in (1, 2, 4, 8):
            for lengths in (None, rng.integers(1000, 6_000_000, n)):
                bands = [sharding.owned_cols(n, r, w, lengths) for r in range(w)]
                assert np.concatenate(bands).tolist() == list(range(n))                 # contiguous, ordered, complete
                if n >= w:
                    assert all(b.size for b in bands)
    # bands cut by bytes: 8 long genomes followed by 504 short ones must not put all the work on rank 0
    lengths = np.array([5_000_000] * 8 + [10_000] * 504)
    b = sharding.band_bounds(lengths, 4)
    w = [lengths[b[r]:b[r + 1]].sum() for r in range(4)]
    assert max(w) <= 15_040_000 and min(w) >= 10_000_000      # (the best contiguous cut of this input has max 15.0 M)


def test_fast_mode_shares_are_balanced_and_complete():
    """--fast-mode: the upper triangle is dealt by whole columns, heaviest first; every (i <= j) job exactly once and
    no rank more than a few percent above the mean (north_star: 'a balanced set of upper-triangle tiles')"""
    rng = np.random.default_rng(1)
    for n, w in ((512, 8), (512, 2), (37, 4), (3, 8)):
        lengths = rng.integers(4_900_000, 5_100_000, n)
        shares = sharding.fast_mode_cols(lengths, w)
        assert sorted(np.concatenate(shares).tolist()) == list(range(n))
        seen = np.zeros((n, n), dtype=np.int32)
        loads = []
        for cols in shares:
            xs, ys = sharding.triangle_jobs(cols)
            assert np.all(xs <= ys)
            seen[xs, ys] += 1
            loads.append(float(lengths[xs].sum() + lengths[ys].sum()))
        assert np.array_equal(seen, np.triu(np.ones((n, n), dtype=np.int32)))
        if n >= 8 * w:
            assert max(loads) <= 1.03 * (sum(loads) / w), (n, w, loads)


def test_lone_carriage_returns_end_lines_like_text_mode(tmp_path):
    """the reference opens FASTA files in text mode (universal newlines): a lone CR ends a line, also a header line"""
    from snacc_b200 import fasta
    cases = [b">a\rACGT\rTTGA\r>b\rGG\r", b">a desc\rAC GT\r\nTT\r\r\nGA\n>b\r\nCC", b"x\r>a\rAC\r\n\rGT\r",
             b">a\nAC\rGT\nTT\r", b">h\r\r\rA\r"]
    for k, t in enumerate(cases):
        f = tmp_path / f"cr{k}.fa"
        f.write_bytes(t)
        want = snacc_oracle.extract_sequences(f, False)
        data, lens = fasta.read_fasta(f)
        assert data.tobytes().decode() == want, (k, t)
        recs = fasta._read_fasta_lines(t)
        assert b"".join(recs).decode() == want and lens == [len(r) for r in recs]


def test_run_log_contents(tmp_path, monkeypatch):
    """A9 (cli.py:147-160): the markdown log names the method, the reverse-complement flag, the output path, the
    versions and every analysed file"""
    import datetime
    files = [tmp_path / "b.fa", tmp_path / "a.fa"]
    txt = gcli.log_template.format(time=datetime.datetime.now(), duration=datetime.timedelta(seconds=3), method="gzip",
                                   rev_comp=True, output_path=tmp_path / "o.csv", py_version="3.12", snacc_version="0.1.0",
                                   jobs=6, pairs=3, compute_s=1.5, csv_s=0.01, gpus=1, pairs_per_s=2.0)
    for needle in ("# `snacc` Analysis", "* Compression method: gzip", "* Reverse complement: True", str(tmp_path / "o.csv"),
                   "## Version Information", "## Analyzed Files", "6 (3 unordered pairs)"):
        assert needle in txt, needle


def test_native_fasta_parser_equals_the_line_parser(tmp_path):
    """read_fasta (snacc_fasta_parse of the C ABI) must give what the line-by-line parser (the reference semantics) gives:
    multi-record files, CRLF, blank lines, trailing blanks, text before the first header, a header at EOF, empty
    records, tabs / form feeds inside and at the end of lines -- and a fuzz over random soups of the characters that
    matter"""
    from snacc_b200 import fasta
    rng = np.random.default_rng(2)
    bodies = []
    for _ in range(6):
        seq = bytes(rng.choice(list(b"ACGTacgtNRY"), size=int(rng.integers(0, 400))).astype(np.uint8))
        lines = [seq[i:i + 60] for i in range(0, len(seq), 60)]
        bodies.append(lines)
    texts = [
        b">a desc\n" + b"\n".join(bodies[0]) + b"\n>b\n" + b"\n".join(bodies[1]) + b"\n",
        b"junk before\n>a\r\n" + b"\r\n".join(bodies[2]) + b"\r\n\r\n>b\r\n" + b"\r\n".join(bodies[3]),
        b">only header",
        b">x\nAC GT  \n\n  ACGT\n>y\n\n>z\nTT\n",
        b"no header at all\nACGT\n",
        b">t\nAC\tGT \t\nAC\x0cGT\n",                    # tabs / form feeds: line parser
        b"",
    ]
    alphabet = np.frombuffer(b"ACGTN>> \t\r\n\n\n\x0b\x0c", dtype=np.uint8)
    for _ in range(60):
        texts.append(bytes(alphabet[rng.integers(0, alphabet.size, int(rng.integers(0, 300)))]))
    many = b"".join(b">r%d\nAC\n" % i for i in range(200))              # more records than the first call makes room for
    texts.append(many)
    for k, t in enumerate(texts):
        f = tmp_path / f"f{k}.fa"
        f.write_bytes(t)
        data, lens = fasta.read_fasta(f)
        data2, lens2 = fasta._read_fasta_native(t)
        recs = fasta._read_fasta_lines(t)
        assert bytes(data) == b"".join(recs) and lens == [len(r) for r in recs], (k, t)
        assert bytes(data2) == bytes(data) and lens2 == lens


def test_direct_csv_writer_is_byte_identical_to_the_pandas_route(tmp_path):
    """write_distance_csv against the reference's own route (cli.py:138-142: DataFrame -> pivot -> to_csv) on labels that
    sort differently as strings and as paths, labels that need quoting, and floats of every shape"""
    import pandas as pd
    rng = np.random.default_rng(1)
    n = 40
    files = [Path(f"/data/g{rng.integers(0, 3)}/sub dir/my,Genome_{i}.fasta") if i % 7 == 0 else
             Path(f"/data/g{rng.integers(0, 3)}/Genome_{i}.fa") for i in range(n)] + [Path("/data/g1"), Path("/data/g1-x/a.fa")]
    n = len(files)
    D = rng.random((n, n))
    D[0, 0], D[0, 1], D[1, 0], D[2, 2], D[3, 3] = 0.0, 1.0, 1e-5, 1.0000000000000002, -0.25
    D[4, 4], D[5, 5], D[6, 6], D[7, 7], D[8, 8] = 1e-300, 0.1 + 0.2, 123456789.125, 1e16, 2.5e-7
    rows = [(files[i], files[j], D[i][j]) for i in range(n) for j in range(n)]
    ref = tmp_path / "ref.csv"
    pd.DataFrame(rows, columns=["file", "file2", "ncd"]).pivot(index="file", columns="file2", values="ncd").to_csv(ref)
    out = tmp_path / "out.csv"
    gcli.write_distance_csv(files, D, out)
    assert out.read_bytes() == ref.read_bytes()


def test_native_csv_writer_formats_every_double_like_repr(tmp_path, monkeypatch):
    """snacc_csv_write (C ABI, host threads) against the Python loop and against pandas: zeros, whole numbers, the
    fixed / exponent switch at 1e-4 and 1e16, subnormals, the largest double, NaN (empty cell), infinities, random
    doubles over 600 binary orders of magnitude"""
    import pandas as pd
    from snacc_b200.engine import load_library
    assert hasattr(load_library(), "snacc_csv_write")
    rng = np.random.default_rng(2)
    n = 60
    D = rng.random((n, n))
    special = [0.0, -0.0, 1.0, 0.5, 1e-5, 1e-4, 0.0001234, 9.999e-5, 1e15, 1e16, 9.999999999999998e15, 1e17, 5e-324, 1e-323,
               2.2250738585072014e-308, 1.7976931348623157e308, float("nan"), float("inf"), -float("inf"), -1.5e-7, 0.1, 1 / 3,
               1e22, 1e21, 100.0, 1234567.0, 0.30000000000000004, 12345678.9, -123456789012345680.0]
    D.flat[:len(special)] = special
    D[5] = np.exp(rng.uniform(-50, 50, n)); D[6] = -np.exp(rng.uniform(-700, 700, n)); D[7] = rng.integers(-10**6, 10**6, n)
    files = [Path(f"/d/g,{i}.fa") if i % 9 == 0 else Path(f"/d/g{i:02d}.fa") for i in range(n)]
    native, plain, ref = tmp_path / "native.csv", tmp_path / "plain.csv", tmp_path / "ref.csv"
    gcli.write_distance_csv(files, D, native)
    monkeypatch.setattr(gcli, "_write_csv_native", lambda *a: False)
    gcli.write_distance_csv(files, D, plain)
    assert native.read_bytes() == plain.read_bytes()
    rows = [(files[i], files[j], D[i][j]) for i in range(n) for j in range(n)]
    pd.DataFrame(rows, columns=["file", "file2", "ncd"]).pivot(index="file", columns="file2", values="ncd").to_csv(ref)
    assert native.read_bytes() == ref.read_bytes()


def test_newick_writer_equals_the_reference_recursion():
    """newick_from_linkage (no scipy) against the reference's to_tree + get_newick recursion on scipy linkages"""
    from oracle import tree_oracle
    from snacc_b200 import distmatrix_to_tree as d2t
    rng = np.random.default_rng(4)
    for n in (2, 3, 9, 150):
        D = tree_oracle.metrify(rng.random((n, n)))
        Z = tree_oracle.hierarchical(D)
        names = [f"g{i}" for i in range(n)]
        assert d2t.newick_from_linkage(Z, names) == tree_oracle.newick(Z, names)

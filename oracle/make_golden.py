"""Generate tests/golden/ from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Writes
  tests/golden/fasta/*.f*a          small deterministic FASTA fixtures (synthetic; seeds below)
  tests/golden/reference_sizes.json ``compressed_size`` / ``compute_distance`` outputs of
                                    /root/reference/snacc/pairwise_ncd.py (imported through
                                    oracle/ref_loader.py) for every fixture, ordered pair,
                                    algorithm in {lz4, gzip, zlib} and reverse_complement flag
  tests/golden/reference_cli_<algo>[_rc].csv  distance CSVs written by the reference's own
                                    ``cli`` (/root/reference/snacc/cli.py:69-142) on the fixtures
  tests/golden/codec_vectors.json   compressed lengths of seeded synthetic byte strings under the
                                    reference's three compressor calls
The reference tree cannot travel to the GPU box, these files can.
"""
import gzip
import json
import os
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _wrap(seq, width=60):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


def _mutate(rng, parent, rate):
    child = parent.copy()
    m = rng.random(child.size) < rate
    child[m] = rng.choice(ACGT, size=int(m.sum()))
    # a few short indels so lengths differ
    for _ in range(3):
        p = int(rng.integers(0, child.size))
        if rng.random() < 0.5:
            child = np.concatenate([child[:p], rng.choice(ACGT, size=int(rng.integers(1, 30))), child[p:]])
        else:
            child = np.concatenate([child[:p], child[p + int(rng.integers(1, 30)):]])
    return child


def make_fixtures():
    fdir = GOLD / "fasta"
    fdir.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(20261018)
    root = rng.choice(ACGT, size=3000)
    for i in range(4):
        seq = _mutate(rng, root, 0.01 * (i + 1)).tobytes().decode()
        (fdir / f"g{i + 1}.fasta").write_text(f">g{i + 1} synthetic relative {i + 1}\n{_wrap(seq)}\n")
    # multi-record file with lower-case, IUPAC codes, a blank line and trailing blanks
    r1 = rng.choice(ACGT, size=400).tobytes().decode()
    r2 = rng.choice(np.frombuffer(b"ACGTacgtNRYKMSWBDHVn", dtype=np.uint8), size=300).tobytes().decode()
    r3 = rng.choice(ACGT, size=150).tobytes().decode().lower()
    (fdir / "multi.fa").write_text(
        f">rec1 first\n{_wrap(r1, 70)}\n\n>rec2 iupac and case  \n{_wrap(r2, 50)}  \n>rec3\n{_wrap(r3, 80)}\n")
    # N run inside otherwise random sequence
    a = rng.choice(ACGT, size=2500)
    a[900:1400] = ord("N")
    (fdir / "nrun.fna").write_text(f">nrun\n{_wrap(a.tobytes().decode())}\n")
    (fdir / "tiny.fa").write_text(">tiny\nACGTA\n")
    # two relatives above the 64 KiB single-block LZ4 regime / deflate window
    big = rng.choice(ACGT, size=70000)
    (fdir / "big1.fasta").write_text(f">big1\n{_wrap(big.tobytes().decode(), 70)}\n")
    (fdir / "big2.fasta").write_text(f">big2\n{_wrap(_mutate(rng, big, 0.03).tobytes().decode(), 70)}\n")
    return sorted(fdir.iterdir(), key=lambda p: str(p.absolute()))


def synth_vector(kind, n, seed):
    """Seeded byte strings shared by make_golden and the tests (tests import this function)."""
    rng = np.random.default_rng(seed)
    if kind == "dna":
        return rng.choice(ACGT, size=n)
    if kind == "bytes":
        return rng.integers(0, 256, size=n, dtype=np.uint8)
    if kind == "run":
        return np.full(n, ord("A"), dtype=np.uint8)
    if kind == "period":
        return np.resize(rng.choice(ACGT, size=int(rng.integers(1, 50))), n)
    if kind == "lower":
        d = rng.choice(ACGT, size=n)
        d[rng.random(n) < 0.3] |= 0x20
        return d
    if kind == "nrun":
        d = rng.choice(ACGT, size=n)
        for _ in range(max(1, n // 20000)):
            s = int(rng.integers(0, max(1, n)))
            d[s:s + int(rng.integers(1, 5000))] = ord("N")
        return d
    if kind == "mut":
        h = n // 2
        x = rng.choice(ACGT, size=h)
        y = x.copy()
        m = rng.random(h) < 0.02
        y[m] = rng.choice(ACGT, size=int(m.sum()))
        return np.concatenate([x, y, rng.choice(ACGT, size=n - 2 * h)])
    if kind == "low":
        return (rng.integers(0, 3, size=n, dtype=np.uint8) + 65).astype(np.uint8)
    raise KeyError(kind)


VECTOR_KINDS = ["dna", "bytes", "run", "period", "lower", "nrun", "mut", "low"]
VECTOR_SIZES = [0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 20, 64, 1000, 11000, 22000, 65273, 65274, 65275, 65535,
                65536, 65537, 65546, 65547, 65548, 98042, 131071, 131072, 131073, 200000]


def main():
    sys.path.insert(0, str(ROOT))
    from oracle import ref_loader
    ref_pkg = ref_loader.load_reference_package()
    ref = sys.modules["_reference_snacc.pairwise_ncd"]
    import importlib
    ref_cli = importlib.import_module("_reference_snacc.cli")
    import lz4framed  # the shim

    files = make_fixtures()
    out = {"bias": sys.getsizeof(b""), "files": [f.name for f in files], "cases": {}}
    for algo in ("lz4", "gzip", "zlib"):
        for rc in (False, True):
            C = [ref.compressed_size(f, algo, reverse_complement=rc)[1] for f in files]
            S = [[ref.compressed_size((a, b), algo, reverse_complement=rc)[1] for b in files] for a in files]
            D = [[ref.compute_distance(C[i], C[j], S[i][j], S[j][i]) for j in range(len(files))]
                 for i in range(len(files))]
            out["cases"][f"{algo}{'_rc' if rc else ''}"] = {"C": C, "S": S, "D": D}
    out["formula_kat"] = {"args": [1174721, 1173133, 1242873, 1242873],
                          "value": ref.compute_distance(1174721, 1173133, 1242873, 1242873)}
    sample = Path(ref_loader.REFERENCE_ROOT) / "test_dataset" / "sample.fa"
    out["sample_fa"] = {}
    for algo in ("lz4", "gzip", "zlib"):
        for rc in (False, True):
            cx = ref.compressed_size(sample, algo, reverse_complement=rc)[1]
            cxx = ref.compressed_size((sample, sample), algo, reverse_complement=rc)[1]
            out["sample_fa"][f"{algo}{'_rc' if rc else ''}"] = {
                "C": cx, "Cxx": cxx, "ncd": ref.compute_distance(cx, cx, cxx, cxx)}
    (GOLD / "reference_sizes.json").write_text(json.dumps(out, indent=1))

    # the reference's own CLI, unmodified, on the small fixtures (big ones excluded to keep CSVs tiny)
    small = [str(f) for f in files if not f.name.startswith("big")]
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for algo in ("lz4", "gzip", "zlib"):
                for rc in (False, True):
                    name = f"reference_cli_{algo}{'_rc' if rc else ''}.csv"
                    ref_cli.cli.callback(sequences=tuple(small), fasta=(), directories=(), numThreads=4,
                                             compression=algo, showProgress=False, saveCompression=None,
                                             output=name, reverse_complement=rc, log=False)
                    text = Path(name).read_text()
                    # header/index cells are absolute paths of this container: keep basenames only
                    text = text.replace(str(GOLD / "fasta") + "/", "")
                    (GOLD / name).write_text(text)
        finally:
            os.chdir(cwd)

    vec = {"sizes": VECTOR_SIZES, "kinds": VECTOR_KINDS, "lz4": {}, "gzip": {}, "zlib": {}}
    for k, kind in enumerate(VECTOR_KINDS):
        for key in ("lz4", "gzip", "zlib"):
            vec[key][kind] = []
        for n in VECTOR_SIZES:
            b = synth_vector(kind, n, 1000 * k + n % 997).tobytes()
            vec["lz4"][kind].append(len(lz4framed.compress(b)))
            vec["gzip"][kind].append(len(gzip.compress(b)))
            vec["zlib"][kind].append(len(zlib.compress(b)))
    (GOLD / "codec_vectors.json").write_text(json.dumps(vec))
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()

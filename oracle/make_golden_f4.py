"""Golden sizes for the two codecs the GPU path does not have yet: lzma -- the reference's default,
cli.py:52 -- and bzip2, produced by the UNMODIFIED reference (pairwise_ncd.py:71-76: lzma.compress, bz2.compress with
their defaults: xz container preset 6, bzip2 level 9) on the committed FASTA fixtures.  Run in the build container only:

    python -m oracle.make_golden_f4        ->  tests/golden/reference_sizes_lzma_bzip2.json

TEST INFRASTRUCTURE ONLY.  These pin the oracle a future device implementation has to meet.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def main():
    sys.path.insert(0, str(ROOT))
    from oracle import ref_loader
    ref_loader.load_reference_package()
    ref = sys.modules["_reference_snacc.pairwise_ncd"]
    files = sorted((GOLD / "fasta").iterdir(), key=lambda p: str(p.absolute()))
    out = {"bias": sys.getsizeof(b""), "files": [f.name for f in files], "cases": {}}
    for algo in ("lzma", "bzip2"):
        for rc in (False, True):
            C = [ref.compressed_size(f, algo, reverse_complement=rc)[1] for f in files]
            S = [[ref.compressed_size((a, b), algo, reverse_complement=rc)[1] for b in files] for a in files]
            D = [[ref.compute_distance(C[i], C[j], S[i][j], S[j][i]) for j in range(len(files))] for i in range(len(files))]
            out["cases"][f"{algo}{'_rc' if rc else ''}"] = {"C": C, "S": S, "D": D}
    (GOLD / "reference_sizes_lzma_bzip2.json").write_text(json.dumps(out, indent=1))
    print("written", GOLD / "reference_sizes_lzma_bzip2.json")


if __name__ == "__main__":
    main()

"""Synthetic mutated-ACGT genome sets of the shapes named in BASELINE.json.

Deterministic: everything derives from ``numpy.random.default_rng(seed)`` (host generator) or a seeded
``torch.Generator`` (device generator used by bench.py for the multi-GB sets).  Genomes form a binary
phylogeny: genome i descends from genome (i-1)//2 through substitutions plus a few short indels, so
leaf-to-leaf divergence spans roughly 0.1 % - 10 % and lengths differ slightly.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutate(rng, parent, sub_rate=0.01, n_indels=4, max_indel=50):
    child = parent.copy()
    m = rng.random(child.size) < sub_rate
    child[m] = rng.choice(ACGT, size=int(m.sum()))
    for _ in range(n_indels):
        p = int(rng.integers(0, max(1, child.size)))
        k = int(rng.integers(1, max_indel + 1))
        if rng.random() < 0.5:
            child = np.concatenate([child[:p], rng.choice(ACGT, size=k), child[p:]])
        else:
            child = np.concatenate([child[:p], child[p + k:]])
    return child


def phylogeny(n_genomes, length, seed, sub_rate=0.01, n_indels=4):
    """List of uint8 arrays (upper-case ACGT)."""
    rng = np.random.default_rng(seed)
    out = [rng.choice(ACGT, size=int(length))]
    for i in range(1, n_genomes):
        out.append(mutate(rng, out[(i - 1) // 2], sub_rate, n_indels))
    return out


def phylogeny_torch(n_genomes, length, seed, device, sub_rate=0.01, n_indels=4, max_indel=50):
    """Same construction on a CUDA device (bench-scale sets); returns a list of uint8 tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    cpu_rng = np.random.default_rng(seed)

    def rnd(n):
        return lut[torch.randint(0, 4, (n,), generator=g, device=device)]

    out = [rnd(int(length))]
    for i in range(1, n_genomes):
        p = out[(i - 1) // 2]
        m = torch.rand(p.numel(), generator=g, device=device) < sub_rate
        c = torch.where(m, rnd(p.numel()), p)
        for _ in range(n_indels):
            pos = int(cpu_rng.integers(0, c.numel()))
            k = int(cpu_rng.integers(1, max_indel + 1))
            if cpu_rng.random() < 0.5:
                c = torch.cat([c[:pos], rnd(k), c[pos:]])
            else:
                c = torch.cat([c[:pos], c[pos + k:]])
        out.append(c)
    return out


def stale_slot_stream(seed):
    """Test input for the LZ4 table encoding: ACGT stretches separated by A/T-only stretches of 66 k - 300 k bases, so
    that every hash-table slot of a k-mer containing C or G goes out of reach -- far more than 131072 positions --
    before it is probed again.  Returns (x, y)."""
    rng = np.random.default_rng(seed)

    def seg(alphabet, n):
        return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]

    parts = []
    for _ in range(8):
        parts.append(seg(b"ACGT", int(rng.integers(60000, 120000))))
        parts.append(seg(b"AT", int(rng.integers(66000, 300000))))
    return seg(b"ACGT", int(rng.integers(1000, 200000))), np.concatenate(parts)


def write_fasta(path, name, seq, width=70):
    s = bytes(seq).decode("ascii")
    with open(path, "w") as fh:
        fh.write(f">{name}\n")
        for i in range(0, len(s), width):
            fh.write(s[i:i + width] + "\n")

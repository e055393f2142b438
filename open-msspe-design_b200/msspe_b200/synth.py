"""Synthetic pre-aligned genome sets of the shapes BASELINE.json names.

ancestor = i.i.d. uniform ACGT; C clade founders = ancestor with per-base substitution probability p_clade;
each genome = a random founder with substitution probability p_leaf; no indels (pre-aligned, equal length);
optional runs of '-' (geometric length, mean 30) starting at rate gap_rate per base, and 'N' at n_rate.
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_genomes(n: int, length: int, seed: int, clades: int = 1, p_clade: float = 0.0, p_leaf: float = 0.02,
                  gap_rate: float = 0.0, n_rate: float = 0.0) -> np.ndarray:
    """Returns an (n, length) uint8 array of ASCII bases."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, length, dtype=np.uint8)
    founders = np.tile(anc, (clades, 1))
    if clades > 1 or p_clade > 0:
        mut = rng.random((clades, length)) < p_clade
        founders[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
    out = np.empty((n, length), dtype=np.uint8)
    which = rng.integers(0, clades, n)
    chunk = max(1, (64 << 20) // max(1, length))
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        g = founders[which[s:e]].copy()
        mut = rng.random((e - s, length)) < p_leaf
        g[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
        a = _ALPHA[g]
        if gap_rate > 0:
            starts = np.argwhere(rng.random((e - s, length)) < gap_rate)
            for r, c in starts:
                ln = int(rng.geometric(1.0 / 30.0))
                a[r, c:c + ln] = ord("-")
        if n_rate > 0:
            a[rng.random((e - s, length)) < n_rate] = ord("N")
        out[s:e] = a
    return out


def to_fasta(genomes: np.ndarray, prefix: str = "g") -> bytes:
    parts = []
    for i in range(genomes.shape[0]):
        parts.append(b">%s%d synthetic\n" % (prefix.encode(), i))
        parts.append(genomes[i].tobytes())
        parts.append(b"\n")
    return b"".join(parts)


def offsets_for(genomes: np.ndarray) -> np.ndarray:
    n, length = genomes.shape
    return (np.arange(n + 1, dtype=np.uint64) * np.uint64(length)).astype(np.uint64)


# The named shapes.  cfg1 is the reference's own CPU-runnable case; cfg2 is the N=1 bench workload.
CONFIGS = {
    "cfg1": dict(n=50, length=10_000, seed=1, clades=1, p_clade=0.0, p_leaf=0.02, k=13),
    "cfg2": dict(n=1_000, length=30_000, seed=2, clades=8, p_clade=0.08, p_leaf=0.01, gap_rate=1e-4, k=13),
    "cfg3": dict(n=10_000, length=11_000, seed=3, clades=64, p_clade=0.10, p_leaf=0.02, k=15),
    "cfg5": dict(n=100_000, length=30_000, seed=5, clades=256, p_clade=0.10, p_leaf=0.01, k=13),
}


def make_config(name: str, scale: float = 1.0) -> tuple[np.ndarray, int]:
    c = dict(CONFIGS[name])
    k = c.pop("k")
    c["n"] = max(2, int(round(c["n"] * scale)))
    return synth_genomes(**c), k


def random_primers(n: int, k: int, seed: int) -> np.ndarray:
    """n distinct uniform-random k-mers as 2-bit codes (cfg4: 20 000 x 13, seed 4)."""
    rng = np.random.default_rng(seed)
    seen = set()
    out = []
    while len(out) < n:
        for v in rng.integers(0, 4 ** k, n, dtype=np.uint64):
            v = int(v)
            if v not in seen:
                seen.add(v)
                out.append(v)
                if len(out) == n:
                    break
    return np.array(out, dtype=np.uint64)

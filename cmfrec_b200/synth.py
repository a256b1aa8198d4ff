"""Synthetic interaction matrices of the shapes BASELINE.json names, fixed seeds.

The real datasets are not available offline; these reproduce their size, density and the heavy-tailed
degree distributions that make the workload hard (a few items with tens of thousands of entries).
Output is COO sorted by (row, col) with distinct pairs, int32 indices.
"""
import numpy as np

SHAPES = {
    # name: (m, n, nnz, seed, kind, mean-degree lognormal sigma)
    "cfg1": (2000, 1000, 20000, 1, "ratings", 0.8),
    "ml10m": (69878, 10677, 10000054, 20260101, "ratings", 1.0),
    "lastfm": (358858, 160112, 17164027, 20260102, "counts", 0.6),
}


ZIPF = {"cfg1": 0.8, "ml10m": 0.62, "lastfm": 0.66}


def _zipf_cols(rng, n, size, exponent):
    w = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** exponent
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side="left").astype(np.int64)


def make(name, dtype=np.float32, scale=1.0):
    """Return (ixA, ixB, X, m, n).  scale < 1 shrinks rows, columns and entries proportionally."""
    m, n, nnz, seed, kind, sigma = SHAPES[name]
    if scale != 1.0:
        m, n, nnz = max(int(m * scale), 8), max(int(n * scale), 8), max(int(nnz * scale), 64)
    rng = np.random.default_rng(seed)
    exponent = ZIPF[name]
    perm = rng.permutation(n)          # popular items are not the low ids
    over = 1.15
    while True:
        # row degrees: lognormal, clipped, rescaled to the (oversampled) target total
        deg = rng.lognormal(0.0, sigma, size=m)
        deg = np.clip(deg / deg.sum() * nnz * over, 1, n // 2)
        deg = np.maximum(np.round(deg / deg.sum() * nnz * over), 1).astype(np.int64)
        rows = np.repeat(np.arange(m, dtype=np.int64), deg)
        cols = perm[_zipf_cols(rng, n, rows.size, exponent)]
        key = np.unique(rows * n + cols)
        if key.size >= nnz:
            break
        over *= 1.25
    if key.size > nnz:
        key = np.sort(rng.choice(key, size=nnz, replace=False))
    rows, cols = key // n, key % n
    if kind == "ratings":
        vals = rng.integers(1, 11, size=rows.size) * 0.5
    else:
        vals = np.ceil(rng.lognormal(1.0, 1.5, size=rows.size))
    return rows.astype(np.int32), cols.astype(np.int32), vals.astype(dtype), m, n

"""Seeded synthetic inputs of the BASELINE.json configs.

The reference's generator (tester.cpp:103-121, 151-167) draws a Bernoulli(1 - sparsity) mask and
U(-1, 1) values from an unseeded mt19937; here the same distribution comes from numpy's PCG64
with explicit seeds so runs are reproducible.  Configs 4 and 5 cannot exist densely and are
generated directly in the CSR(A^T) form `Plan.from_csc` takes.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (M, N, weight sparsity, activation sparsity)
    "c0": (4096, 4096, 0.5, 0.5),       # what test/main.cpp really runs (tester.cpp:106,154)
    "c1": (4096, 4096, 0.9, 0.5),
    "c2": (4096, 14336, 0.7, 0.5),      # LLM decode FFN up-proj
    "c3": (14336, 4096, 0.7, 0.9),      # down-proj
}


def gen_matrix(M, N, sparsity, seed=1234):
    rng = np.random.default_rng(seed)
    keep = rng.random((M, N), dtype=np.float32) > sparsity
    vals = rng.uniform(-1.0, 1.0, (M, N)).astype(np.float32)
    vals[~keep] = 0.0
    return vals


def gen_vector(M, sparsity, seed=4321):
    rng = np.random.default_rng(seed)
    keep = rng.random(M) > sparsity
    vals = rng.uniform(-1.0, 1.0, M).astype(np.float32)
    return np.where(keep, vals, np.float32(0.0)).astype(np.float32)


def bernoulli_csc(M, N, density, seed):
    """Every cell non-zero with probability `density` (column length ~ Binomial(M, density)),
    built without the dense matrix: geometric gaps over the column-major cell index."""
    rng = np.random.default_rng(seed)
    total = M * N
    expect = int(total * density)
    n_draw = int(expect + 6 * np.sqrt(max(expect, 1)) + 64)
    pos = np.cumsum(rng.geometric(density, n_draw).astype(np.int64)) - 1
    while pos[-1] < total:                      # rare: extend
        more = np.cumsum(rng.geometric(density, n_draw // 8 + 64).astype(np.int64)) + pos[-1]
        pos = np.concatenate([pos, more])
    pos = pos[: np.searchsorted(pos, total)]
    col = pos // M
    row_idx = (pos - col * M).astype(np.int32)
    col_ptr = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(col, minlength=N), out=col_ptr[1:])
    vals = rng.uniform(-1.0, 1.0, pos.size).astype(np.float32)
    vals[vals == 0.0] = 0.5
    return col_ptr, row_idx, vals


def powerlaw_csc(M, N, seed=42, scale=8.0, cap=None):
    """Config 4: column length min(floor(scale * (1-u)^(-1/2)), cap) (Pareto alpha = 2, mean
    about 2*scale), row ids uniform, sorted, distinct."""
    rng = np.random.default_rng(seed)
    cap = M if cap is None else cap
    ln = np.minimum(np.floor(scale * (1.0 - rng.random(N)) ** -0.5), cap).astype(np.int64)
    col_ptr = np.zeros(N + 1, np.int64)
    np.cumsum(ln, out=col_ptr[1:])
    nnz = int(col_ptr[-1])
    col = np.repeat(np.arange(N, dtype=np.int64), ln)
    row = rng.integers(0, M, nnz, dtype=np.int64)
    key = np.unique(col * M + row)              # sorted by (col, row), duplicates dropped
    col = key // M
    row_idx = (key - col * M).astype(np.int32)
    col_ptr = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(col, minlength=N), out=col_ptr[1:])
    vals = rng.uniform(-1.0, 1.0, key.size).astype(np.float32)
    vals[vals == 0.0] = 0.5
    return col_ptr, row_idx, vals

"""Seeded random shapes / sparsities / options for every variant against the oracle: exercises the
corners of the work decompositions (pieces spanning several slabs, more CTAs than rows, single
warps, row panels, multi-row chunks, forced geometries)."""
import numpy as np
import pytest

import oracle_bindings as ob

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(2026)
    out = []
    for k in range(40):
        M = int(rng.choice([1, 7, 31, 32, 33, 100, 257, 1000, 2049, 5000]))
        N = int(rng.choice([32, 64, 288, 1024, 4128, 8192]))
        sa = float(rng.choice([0.0, 0.5, 0.9, 0.99, 0.999]))
        sx = float(rng.choice([0.0, 0.5, 0.95]))
        opts = {}
        if rng.random() < 0.6:
            opts["row_splits"] = int(rng.choice([1, 2, 5, 17, 64]))
        if rng.random() < 0.5:
            opts["warps_per_col"] = int(rng.choice([1, 2, 4, 8]))
        if rng.random() < 0.5:
            opts["slab_cols"] = int(rng.choice([256, 512, 2048, 4096]))
        if rng.random() < 0.5:
            opts["chunk_mode"] = int(rng.choice([1, 2]))
        out.append((k, M, N, sa, sx, opts))
    return out


@pytest.mark.parametrize("k,M,N,sa,sx,opts", _cases())
def test_random_case(k, M, N, sa, sx, opts):
    import spmv_test_b200 as S
    A = ob.gen_matrix(M, N, sa, 7000 + k)
    x = ob.gen_vector(M, sx, 8000 + k)
    ptr, idx, val = ob.dense_to_csc(A)
    y_ref = ob.csc_gemv(N, ptr, idx, val, x)
    s = ob.csc_gemv(N, ptr, idx, np.abs(val), np.abs(x)).astype(np.float64)
    for v in ("wsp", "asp", "awsp", "tcsr"):
        o = dict(opts)
        if v == "wsp":
            o = {kk: vv for kk, vv in o.items() if kk == "warps_per_col"}
        if v == "asp":
            o = {kk: vv for kk, vv in o.items() if kk == "row_splits"}
        with S.Plan.from_dense(v, A, **o) as p:
            y = p.run_host(x)
            err = np.abs(y.astype(np.float64) - y_ref)
            assert float(np.max(err / (s + 1e-30), initial=0.0)) <= 1e-5, (v, M, N, sa, sx, o)
            assert float(np.max(err, initial=0.0)) <= 1e-3
            assert p.run_host(x).tobytes() == y.tobytes(), (v, "non-deterministic")


def _lob_cases():
    rng = np.random.default_rng(3033)
    out = []
    for k in range(16):
        M = int(rng.choice([1, 33, 1000, 1024, 1025, 3000, 9000]))
        N = int(rng.choice([32, 1024, 2080, 4128, 8192]))
        sa = float(rng.choice([0.5, 0.95, 0.99, 0.999]))
        sx = float(rng.choice([0.0, 0.5, 0.95]))
        opts = {"chunk_mode": 3}
        if rng.random() < 0.5:
            opts["row_splits"] = int(rng.choice([1, 2, 5]))
        if rng.random() < 0.6:
            opts["slab_cols"] = int(rng.choice([1024, 2048, 4096]))
        out.append((k, M, N, sa, sx, opts))
    return out


@pytest.mark.parametrize("k,M,N,sa,sx,opts", _lob_cases())
def test_random_case_lane_owned_blocks(k, M, N, sa, sx, opts):
    import spmv_test_b200 as S
    A = ob.gen_matrix(M, N, sa, 9000 + k)
    x = ob.gen_vector(M, sx, 9500 + k)
    ptr, idx, val = ob.dense_to_csc(A)
    y_ref = ob.csc_gemv(N, ptr, idx, val, x)
    s = ob.csc_gemv(N, ptr, idx, np.abs(val), np.abs(x)).astype(np.float64)
    for v in ("awsp", "tcsr"):
        with S.Plan.from_dense(v, A, **opts) as p:
            y = p.run_host(x)
            err = np.abs(y.astype(np.float64) - y_ref)
            assert float(np.max(err / (s + 1e-30), initial=0.0)) <= 1e-5, (v, M, N, sa, sx, opts)
            assert float(np.max(err, initial=0.0)) <= 1e-3
            assert p.run_host(x).tobytes() == y.tobytes(), (v, "non-deterministic")

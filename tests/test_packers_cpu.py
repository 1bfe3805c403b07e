"""CPU tests of the product's device-format packers through spmv_pack_dump_* (host image of
what would be uploaded).  The decode below is numpy test code; the product has no CPU SGEMV."""
import numpy as np
import pytest

import oracle_bindings as ob


def decode_panel(d):
    """Rebuilds the dense matrix from a dumped AWSP / TCSR image."""
    A = np.zeros((d.M, d.N), np.float32)
    W = d.slab_cols
    vals = d.vals.reshape(-1, 4)
    idx = d.idx.reshape(-1, 4).astype(np.int64)
    tiled = d.variant == "tcsr"
    for s in range(d.slabs):
        for row in range(d.M):
            if tiled:
                rb, r = divmod(row, 32)
                tb = int(d.off[s * (d.row_blocks + 1) + rb])
                te = int(d.off[s * (d.row_blocks + 1) + rb + 1])
                rel = d.rel[(s * d.row_blocks + rb) * 32:(s * d.row_blocks + rb + 1) * 32].astype(np.int64)
                g0 = tb + rel[r]
                g1 = tb + rel[r + 1] if r < 31 else te
            else:
                g0 = int(d.off[s * (d.M + 1) + row])
                g1 = int(d.off[s * (d.M + 1) + row + 1])
            if g1 == g0:
                continue
            v = vals[g0:g1].reshape(-1)
            c = idx[g0:g1].reshape(-1)
            real = v != 0
            # pads (value 0) use one column that is absent from the row; real columns are distinct
            assert not (set(c[~real]) & set(c[real])) and len(set(c[~real])) <= 1
            assert np.all(c < W)
            assert len(set(c[real])) == int(real.sum()), "a column occurs once per segment"
            # bank-aware order: within a chunk of 32 groups, element e of the lanes hits distinct
            # banks unless some bank holds more than four of the row's columns
            for q0 in range(0, g1 - g0, 32):
                blk = idx[g0 + q0:g0 + min(q0 + 32, g1 - g0)]
                worst = max(np.bincount(blk[:, e] % 32).max() for e in range(4))
                need = -(-np.bincount(blk.reshape(-1) % 32).max() // 4)
                assert worst <= max(need, 1) + 1
            A[row, s * W + c[real]] = v[real]
    return A


def decode_wsp(d):
    panels, prow = d.slabs, d.slab_cols            # wsp dump: row panels, rows per panel (pad index)
    A = np.zeros((max(d.M, panels * prow) + 1, d.N), np.float32)
    vals = d.vals.reshape(-1, 4)
    idx = d.idx.reshape(-1, 4).astype(np.int64)
    assert vals.shape[0] == d.groups + 1 and np.all(vals[-1] == 0) and np.all(idx[-1] == prow)   # spare pad group
    assert d.off.size == panels * d.N + 1
    for p in range(panels):
        for c in range(d.N):
            g0, g1 = int(d.off[p * d.N + c]), int(d.off[p * d.N + c + 1])
            v = vals[g0:g1].reshape(-1)
            r = idx[g0:g1].reshape(-1)
            assert np.all(r[v == 0] == prow) and np.all(v[r == prow] == 0)
            real = v != 0
            assert len(set(r[real])) == int(real.sum()), "a row occurs once per list"
            assert np.all(r[real] < prow)
            A[p * prow + r[real], c] = v[real]
    return A[:d.M]


CASES = [(64, 32, 0.5), (96, 512, 0.9), (1000, 288, 0.7), (33, 1024, 0.3), (256, 256, 0.0), (128, 64, 1.0)]


@pytest.mark.parametrize("M,N,sa", CASES)
@pytest.mark.parametrize("variant", ["awsp", "tcsr"])
def test_panel_roundtrip(M, N, sa, variant):
    import spmv_test_b200 as S
    A = ob.gen_matrix(M, N, sa, M * 7 + N)
    d = S.pack_dump(variant, A)
    assert d.nnz == np.count_nonzero(A) and d.vals.size == 4 * d.groups == d.idx.size
    assert np.array_equal(decode_panel(d), A)
    # CSR(A^T) input gives the identical image
    ptr, idx, val = ob.dense_to_csc(A)
    e = S.pack_dump(variant, csc=(ptr, idx, val), shape=(M, N))
    for f in ("vals", "idx", "off", "rel"):
        assert getattr(d, f).tobytes() == getattr(e, f).tobytes(), f


@pytest.mark.parametrize("slab", [256, 512, 1024, 4096])
def test_panel_slab_widths(slab):
    import spmv_test_b200 as S
    A = ob.gen_matrix(160, 4096 + 512, 0.95, slab)
    for variant in ("awsp", "tcsr"):
        d = S.pack_dump(variant, A, slab_cols=slab)
        assert d.slab_cols == slab and d.index_bits == (8 if slab == 256 else 16)
        assert np.array_equal(decode_panel(d), A)


def test_panel_slab_auto_from_density():
    import spmv_test_b200 as S
    assert S.pack_dump("awsp", ob.gen_matrix(2048, 2048, 0.7, 1)).slab_cols == 256
    wide = S.pack_dump("awsp", ob.gen_matrix(4096, 8192, 0.99, 2))      # few slabs / work units: narrowed
    assert wide.slab_cols == 512 and wide.index_bits == 16
    from spmv_test_b200 import synth
    cp, ri, va = synth.bernoulli_csc(32768, 65536, 0.01, 9)               # config-5 family: short segments even at
    assert S.pack_dump("awsp", csc=(cp, ri, va), shape=(32768, 65536)).slab_cols == 2048   # 4096 -> 2048 (two CTAs per SM)


@pytest.mark.parametrize("M,N,sa", CASES)
def test_wsp_roundtrip(M, N, sa):
    import spmv_test_b200 as S
    A = ob.gen_matrix(M, N, sa, M * 5 + N)
    d = S.pack_dump("wsp", A)
    assert d.index_bits == 16 and d.nnz == np.count_nonzero(A) and d.vals.size == 4 * (d.groups + 1)
    assert np.array_equal(decode_wsp(d), A)
    d32 = S.pack_dump("wsp", A, index_bits=32)
    assert d32.index_bits == 32 and np.array_equal(decode_wsp(d32), A)
    ptr, idx, val = ob.dense_to_csc(A)
    e = S.pack_dump("wsp", csc=(ptr, idx, val), shape=(M, N))
    assert d.vals.tobytes() == e.vals.tobytes() and d.idx.tobytes() == e.idx.tobytes() and d.off.tobytes() == e.off.tobytes()


def test_rejects_bad_input():
    import spmv_test_b200 as S
    with pytest.raises(S.SpmvError):
        S.pack_dump("awsp", np.zeros((32, 40), np.float32))
    with pytest.raises(S.SpmvError):
        S.pack_dump("asp", np.zeros((32, 32), np.float32))
    ptr = np.array([0, 1] + [1] * 31, np.int64)
    with pytest.raises(S.SpmvError):      # row index out of range
        S.pack_dump("awsp", csc=(ptr, np.array([99], np.int32), np.array([1.0], np.float32)), shape=(32, 32))


def test_synthetic_generators():
    from spmv_test_b200 import synth
    cp, ri, va = synth.bernoulli_csc(4096, 512, 0.01, 3)
    assert cp[-1] == ri.size == va.size and abs(ri.size / (4096 * 512) - 0.01) < 0.001
    for c in (0, 17, 511):
        seg = ri[cp[c]:cp[c + 1]]
        assert np.all(np.diff(seg) > 0) and (seg.size == 0 or (seg.min() >= 0 and seg.max() < 4096))
    cp, ri, va = synth.powerlaw_csc(1 << 14, 1 << 14, seed=1)
    ln = np.diff(cp)
    assert 10 < ln.mean() < 24 and ln.max() > 20 * ln.mean()
    assert np.all(va != 0)


def test_wsp_row_panels_for_tall_matrices():
    """x of a tall matrix does not fit shared memory: rows are cut into 16384-row panels (12288 rows) with
    panel-local 16-bit row ids when the lists stay long enough."""
    import spmv_test_b200 as S
    A = ob.gen_matrix(40000, 64, 0.99, 17)
    d = S.pack_dump("wsp", A)
    assert d.slabs == 4 and d.slab_cols == 12288 and d.index_bits == 16
    assert np.array_equal(decode_wsp(d), A)
    ptr, idx, val = ob.dense_to_csc(A)
    e = S.pack_dump("wsp", csc=(ptr, idx, val), shape=A.shape)
    assert d.vals.tobytes() == e.vals.tobytes() and d.idx.tobytes() == e.idx.tobytes() and d.off.tobytes() == e.off.tobytes()
    thin = S.pack_dump("wsp", ob.gen_matrix(40000, 64, 0.9995, 18))      # lists too short: one panel, x through L2
    assert thin.slabs == 1 and thin.slab_cols == 40000


# ---- lane-owned blocks (chunk_mode 3) ---------------------------------------------------------------
def _decode_lane_owned(d):
    """Dense matrix from the lane-owned block format (formats.hpp), checking its invariants."""
    W, R = d.slab_cols, d.block_rows
    cbits = (W // 32).bit_length() - 1
    assert R == min(1024, 1 << (16 - cbits)) and d.index_bits == 16 and d.groups % 32 == 0
    nb = (d.M + R - 1) // R
    A = np.zeros((d.M, d.N), np.float32)
    seen = np.zeros((d.M, d.N), bool)
    off = d.off.reshape(d.slabs, nb + 1)
    vals = d.vals.reshape(-1, 4)
    idx = d.idx.reshape(-1, 4)
    for s in range(d.slabs):
        for b in range(nb):
            g0, g1 = int(off[s, b]), int(off[s, b + 1])
            assert g0 % 32 == 0 and g1 % 32 == 0 and g0 <= g1
            for g in range(g0, g1):
                lane = (g - g0) % 32
                for e in range(4):
                    v, i = vals[g, e], int(idx[g, e])
                    row, col = b * R + (i >> cbits), s * W + (i & ((1 << cbits) - 1)) * 32 + lane
                    assert row < d.M                                   # pads too: x[row] must exist
                    if v != 0.0 or np.isnan(v):
                        assert col < d.N and not seen[row, col]
                        seen[row, col] = True
                        A[row, col] = v
            if b + 1 < nb:
                assert off[s, b + 1] >= off[s, b]
        if s + 1 < d.slabs:
            assert off[s + 1, 0] == off[s, nb]
    return A


@pytest.mark.parametrize("shape,keep,W", [((100, 256), 0.3, 1024), ((1500, 4096), 0.02, 0), ((2100, 2048 + 64), 0.01, 2048),
                                          ((64, 1024), 1.0, 1024), ((40, 512), 0.0, 4096)])
def test_lane_owned_blocks_hold_the_matrix(shape, keep, W):
    import spmv_test_b200 as S
    rng = np.random.default_rng(shape[0] + shape[1])
    A = rng.uniform(-1, 1, shape).astype(np.float32)
    A[rng.random(shape) >= keep] = 0.0
    if keep > 0:
        A[:, 7] = 1.5                                                   # a dense column: one lane's stream is much longer
    kw = {"slab_cols": W} if W else {}
    d = S.pack_dump("awsp", A, chunk_mode=3, **kw)
    assert d.block_rows == min(1024, (65536 * 32) // d.slab_cols) and d.slab_cols == (W or 2048)
    assert d.nnz == np.count_nonzero(A)
    assert np.array_equal(_decode_lane_owned(d), A)
    # the CSR(A^T) source builds the same bytes
    from scipy import sparse
    c = sparse.csc_matrix(A)
    d2 = S.pack_dump("awsp", csc=(c.indptr, c.indices, c.data), shape=shape, chunk_mode=3, **kw)
    assert d2.vals.tobytes() == d.vals.tobytes() and d2.idx.tobytes() == d.idx.tobytes() and d2.off.tobytes() == d.off.tobytes()


def test_lane_owned_blocks_reject_narrow_slabs():
    import spmv_test_b200 as S
    A = ob.gen_matrix(64, 512, 0.5, 3)
    for w in (256, 512):
        with pytest.raises(S.SpmvError):
            S.pack_dump("awsp", A, chunk_mode=3, slab_cols=w)


# ---- row strips (chunk_mode 4) -------------------------------------------------------------------
def _decode_strips(d):
    """Rebuilds the dense matrix from a dumped row-strip image and checks the format's invariants."""
    S16, sw, bands = d.row_blocks, d.slab_cols, d.slabs
    assert d.block_rows == -1 and S16 == 16 and d.index_bits == 32 and sw % 32 == 0
    assert bands == max(1, -(-d.N // (S16 * sw)))
    assert d.off.size == bands * d.M * S16 + 1 and d.off[0] == 0 and d.off[-1] == d.vals.size == d.idx.size
    assert np.all(np.diff(d.off.astype(np.int64)) >= 0) and np.all(d.off % 4 == 0), "segments start on 32-byte boundaries"
    pad = d.idx == 0xFFFFFFFF
    assert np.all(d.vals[pad] == 0) and int((~pad).sum()) == d.nnz
    A = np.zeros((d.M, bands * S16 * sw), np.float32)
    seg = 0
    for b in range(bands):
        for row in range(d.M):
            for s in range(S16):
                e0, e1 = int(d.off[seg]), int(d.off[seg + 1])
                seg += 1
                real = ~pad[e0:e1]
                assert e1 - e0 - int(real.sum()) < 4 and np.all(real[:int(real.sum())]), "fewer than 4 pads, all at the end"
                c = d.idx[e0:e1][real].astype(np.int64)
                assert np.all(c < sw) and np.all(np.diff(c) > 0), "columns of a segment ascend strictly"
                A[row, (b * S16 + s) * sw + c] = d.vals[e0:e1][real]
    assert not np.any(A[:, d.N:])
    return A[:, :d.N]


@pytest.mark.parametrize("shape,keep,sw", [((64, 64), 0.5, 32), ((100, 4096), 0.02, 0), ((37, 2048 + 96), 0.05, 64),
                                           ((300, 1024), 0.3, 32), ((1, 32), 1.0, 0), ((0, 64), 0.5, 0)])
def test_row_strips_hold_the_matrix(shape, keep, sw):
    import spmv_test_b200 as S
    from scipy import sparse
    M, N = shape
    A = ob.gen_matrix(M, N, 1.0 - keep, 99)
    if M > 10:
        A[5, :] = 0.0                                     # an empty row
        A[7, :] = 1.5                                     # a full row: segments longer than a 32-lane window
    kw = {"slab_cols": sw} if sw else {}
    d = S.pack_dump("awsp", A, chunk_mode=4, **kw)
    assert np.array_equal(_decode_strips(d), A)
    assert d.nnz == int(np.count_nonzero(A))
    if M:
        c = sparse.csc_matrix(A)
        d2 = S.pack_dump("awsp", csc=(c.indptr.astype(np.int64), c.indices.astype(np.int32), c.data.astype(np.float32)),
                         shape=(M, N), chunk_mode=4, **kw)
        for f in ("vals", "idx", "off"):
            assert getattr(d, f).tobytes() == getattr(d2, f).tobytes(), f
        assert d2.slab_cols == d.slab_cols


def test_row_strips_width_rule_and_bad_input():
    import spmv_test_b200 as S
    from spmv_test_b200 import synth
    # config-5-like density: ~20.5 non-zeros per (row, strip), whole bands over N
    cp, ri, va = synth.bernoulli_csc(4096, 32768, 0.01, 5)
    d = S.pack_dump("awsp", csc=(cp, ri, va), shape=(4096, 32768), chunk_mode=4)
    assert d.slab_cols == 2048 and d.slabs == 1
    per_seg = np.diff(d.off.astype(np.int64))
    assert 19.5 < d.nnz / per_seg.size < 21.5 and (per_seg > 32).mean() < 0.01
    with pytest.raises(S.SpmvError):                      # strip width must be a multiple of 32 up to the shared-memory limit
        S.pack_dump("awsp", csc=(cp, ri, va), shape=(4096, 32768), chunk_mode=4, slab_cols=48)
    with pytest.raises(S.SpmvError):
        S.pack_dump("awsp", csc=(cp, ri, va), shape=(4096, 32768), chunk_mode=4, slab_cols=4096)
    # a repeated (row, column) pair would put two entries of one window on one accumulator: rejected
    cp2 = np.array([0, 2] + [2] * 31, np.int64)
    with pytest.raises(S.SpmvError):
        S.pack_dump("awsp", csc=(cp2, np.array([3, 3], np.int32), np.array([1.0, 2.0], np.float32)), shape=(8, 32), chunk_mode=4)


@pytest.mark.parametrize("variant,opts", [("wsp", {}), ("awsp", {}), ("tcsr", {}), ("awsp", {"chunk_mode": 3, "slab_cols": 1024}),
                                          ("awsp", {"chunk_mode": 4})])
def test_csc_input_rule_repeated_or_unsorted_entries_are_rejected(variant, opts):
    """a repeated (row, column) pair used to pack silently (awsp/tcsr lost an update, wsp summed)."""
    import spmv_test_b200 as S
    cp = np.array([0, 2] + [2] * 31, np.int64)
    va = np.array([1.0, 2.0], np.float32)
    for rows in ([0, 0], [5, 3], [1, 8]):                 # repeated, descending, out of range (M = 8)
        with pytest.raises(S.SpmvError):
            S.pack_dump(variant, csc=(cp, np.array(rows, np.int32), va), shape=(8, 32), **opts)
    d = S.pack_dump(variant, csc=(cp, np.array([3, 5], np.int32), va), shape=(8, 32), **opts)
    assert d.nnz == 2

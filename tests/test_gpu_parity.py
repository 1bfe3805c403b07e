"""GPU parity tests: the sm_100a kernels, called through the C-ABI, against the CPU oracle
(oracle/spmv_oracle.c, a restatement of reference tester.cpp:36-45 and of the reference
kernels' decodes) on the same seeded inputs."""
import numpy as np
import pytest

import oracle_bindings as ob
from parity import check_y

pytestmark = pytest.mark.gpu

VARIANTS = ["wsp", "asp", "awsp", "tcsr"]


@pytest.fixture(scope="module")
def S():
    import spmv_test_b200 as s
    assert s.lib().spmv_device_count() > 0, "no CUDA device visible to libspmv_b200.so"
    return s


def refs(A, x):
    y32 = ob.sgemv_dense(A, x)
    y64, s = ob.sgemv_dense_f64(A, x)
    return y32, y64, s


def run_all(S, A, x, variants=VARIANTS, **opts):
    y32, y64, s = refs(A, x)
    out = {}
    for v in variants:
        with S.Plan.from_dense(v, A, **opts) as p:
            y = p.run_host(x)
            check_y(y, y32, y64, s, f"{v} {A.shape} {opts}")
            y2 = p.run_host(x)
            assert y.tobytes() == y2.tobytes(), f"{v}: two runs differ (non-deterministic)"
            out[v] = y
    return out


SHAPES = [
    (32, 32, 0.5, 0.5), (64, 32, 0.0, 0.0), (96, 64, 0.9, 0.5), (512, 256, 0.7, 0.5),
    (1000, 512, 0.7, 0.5), (4096, 288, 0.5, 0.9), (33, 1024, 0.3, 0.0), (2048, 2048, 0.99, 0.5),
    (1024, 128, 0.5, 0.5),
]


@pytest.mark.parametrize("M,N,sa,sx", SHAPES)
def test_small_shapes(S, M, N, sa, sx):
    A = ob.gen_matrix(M, N, sa, 1234)
    x = ob.gen_vector(M, sx, 4321)
    run_all(S, A, x)


def test_config0_reference_harness_shape(S):
    """test/main.cpp: 4096x4096, 50 % / 50 % (tester.cpp:106,154)."""
    A = ob.gen_matrix(4096, 4096, 0.5, 1234)
    x = ob.gen_vector(4096, 0.5, 4321)
    ys = run_all(S, A, x)
    # the oracle's restatements of the reference kernels' own decodes agree with the same gate
    y32, y64, s = refs(A, x)
    for layout, ver in (("awsp_ref", 0), ("awsp", 2), ("asp", 2), ("wsp", 0)):
        check_y(ob.decode_gemv(layout, A, x, gpu_order=1, version=ver), y32, y64, s, f"oracle {layout}")
    assert set(ys) == set(VARIANTS)


def test_config1_4096_90pct(S):
    A = ob.gen_matrix(4096, 4096, 0.9, 1234)
    x = ob.gen_vector(4096, 0.5, 4321)
    run_all(S, A, x)


def test_config2_ffn_up(S):
    A = ob.gen_matrix(4096, 14336, 0.7, 1234)
    x = ob.gen_vector(4096, 0.5, 4321)
    run_all(S, A, x)


def test_config3_ffn_down(S):
    A = ob.gen_matrix(14336, 4096, 0.7, 1234)
    x = ob.gen_vector(14336, 0.9, 4321)
    run_all(S, A, x)


def test_edge_zero_matrix_and_zero_x(S):
    A = np.zeros((256, 128), np.float32)
    x = ob.gen_vector(256, 0.5, 1)
    for v in VARIANTS:
        with S.Plan.from_dense(v, A) as p:
            assert not p.run_host(x).any()
    A = ob.gen_matrix(256, 128, 0.5, 2)
    for v in VARIANTS:
        with S.Plan.from_dense(v, A) as p:
            assert not p.run_host(np.zeros(256, np.float32)).any()


def test_edge_negative_zero_and_single_entries(S):
    A = np.zeros((128, 64), np.float32)
    A[5, 7] = 2.0
    A[127, 63] = -3.0
    A[0, 0] = -0.0            # -0.0f == 0.0f: dropped by every packer (matrix_csr.cpp:15)
    x = np.zeros(128, np.float32)
    x[5] = 0.5
    x[127] = 4.0
    x[3] = -0.0               # -0.0f != 0.0f is false: inactive (asp.cu:23)
    out = run_all(S, A, x)
    for y in out.values():
        assert y[7] == 1.0 and y[63] == -12.0 and np.count_nonzero(y) == 2


def test_edge_full_density_and_denormals(S):
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, (160, 96)).astype(np.float32)
    A[A == 0] = 0.5
    A[3, :] = np.float32(1e-40)        # denormal weights are non-zero: kept
    x = rng.uniform(-1, 1, 160).astype(np.float32)
    x[9] = np.float32(1e-41)
    run_all(S, A, x)


def test_empty_shapes(S):
    for M, N in ((0, 64), (64, 0), (0, 0)):
        A = np.zeros((M, N), np.float32)
        for v in VARIANTS:
            with S.Plan.from_dense(v, A) as p:
                y = p.run_host(np.ones(M, np.float32))
                assert y.shape == (N,) and not y.any()


def test_shape_rejected(S):
    with pytest.raises(S.SpmvError) as e:
        S.Plan.from_dense("wsp", np.zeros((64, 48), np.float32))
    assert e.value.code == -2


def test_column_slab_view_with_lda(S):
    """The multi-GPU partitioner packs a column slab of a wider matrix in place (lda = N_total)."""
    A = ob.gen_matrix(512, 1024, 0.7, 5)
    x = ob.gen_vector(512, 0.5, 6)
    y32, y64, s = refs(A, x)
    for v in VARIANTS:
        for a, b in ((0, 256), (256, 1024), (768, 1024)):
            with S.Plan.from_dense(v, A[:, a:b]) as p:
                check_y(p.run_host(x), y32[a:b], y64[a:b], s[a:b], f"{v} slab {a}:{b}")


def test_asp_register_path_long_and_short_lists(S):
    """asp streams chunks with >= 48 active rows with the rows in flight in registers and shorter lists through the
    cp.async ring (asp.cu, kAspRegsMin), in ONE kernel: a CTA whose first 1024-row chunk is long and whose second is
    short, a column tile that ends inside a 512-column tile, a slab view (lda > N), list lengths on both sides of the
    threshold and of the pipeline depth, and lengths that are not a multiple of the four-row groups."""
    M, N = 2048, 992                                         # (N must be a multiple of 32: tester.cpp:9-10)
    wide = ob.gen_matrix(M, N + 32, 0.3, 77)
    A = wide[:, 16:16 + N]                                   # lda = N + 32
    for active_lo, active_hi in ((1024, 3), (96, 0), (95, 95), (513, 97), (0, 1024), (48, 47), (49, 0), (1, 50)):
        x = np.zeros(M, np.float32)
        rng = np.random.default_rng(active_lo * 7 + active_hi)
        x[rng.choice(1024, active_lo, replace=False)] = rng.uniform(-1, 1, active_lo).astype(np.float32)
        x[1024 + rng.choice(1024, active_hi, replace=False)] = rng.uniform(-1, 1, active_hi).astype(np.float32)
        y32, y64, s = refs(np.ascontiguousarray(A), x)
        for splits in (1, 2, 5):
            with S.Plan.from_dense("asp", A, row_splits=splits) as p:
                y = p.run_host(x)
                check_y(y, y32, y64, s, f"asp registers/ring {active_lo}+{active_hi} splits {splits}")
                assert p.run_host(x).tobytes() == y.tobytes()


_TMA_CHILD = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, {here!r}); sys.path.insert(0, {root!r})
import oracle_bindings as ob
from parity import check_y
import spmv_test_b200 as S
M, N = 2048, 992
wide = ob.gen_matrix(M, N + 32, 0.3, 77)
A = wide[:, 16:16 + N]
for seed, sx in ((1, 0.5), (2, 0.97), (3, 0.0)):
    x = ob.gen_vector(M, sx, 900 + seed)
    y32 = ob.sgemv_dense(np.ascontiguousarray(A), x)
    y64, s = ob.sgemv_dense_f64(np.ascontiguousarray(A), x)
    for splits in (1, 3, 0):
        with S.Plan.from_dense("asp", A, row_splits=splits) as p:
            y = p.run_host(x)
            check_y(y, y32, y64, s, "asp tma")
            print(seed, splits, hashlib.sha256(y.tobytes()).hexdigest())
"""


def test_asp_tma_row_gather_path(S):
    """asp_tma_kernel (cp.async.bulk.tensor ... tile::gather4, selected with SPMV_ASP_TMA=1 — read once per process, hence
    the child process): inside the parity gate against the oracle and BIT-IDENTICAL to the default kernel (same FMAs in
    the same order) on a slab view (lda > N), a last tile that ends inside its second 256-column box, dense / half /
    3 % active x, one, three and the default number of row splits."""
    import hashlib
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    code = _TMA_CHILD.format(here=here, root=root)

    def run(env_val):
        env = dict(os.environ, SPMV_ASP_TMA=env_val)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout.strip().splitlines()

    with_tma, without = run("1"), run("0")
    assert len(with_tma) == 9 and with_tma == without


@pytest.mark.parametrize("opts", [dict(row_splits=1), dict(row_splits=3), dict(row_splits=64),
                                  dict(slab_cols=512), dict(slab_cols=4096), dict(index_bits=32),
                                  dict(warps_per_col=1), dict(warps_per_col=8), dict(chunk_mode=1), dict(chunk_mode=2),
                                  dict(chunk_mode=2, slab_cols=256, row_splits=5), dict(chunk_mode=2, warps_per_col=2)])
def test_options(S, opts):
    A = ob.gen_matrix(2048, 1024, 0.8, 11)
    x = ob.gen_vector(2048, 0.5, 12)
    run_all(S, A, x, **opts)


def test_csc_construction_matches_dense(S):
    A = ob.gen_matrix(1024, 512, 0.9, 21)
    x = ob.gen_vector(1024, 0.5, 22)
    ptr, idx, val = ob.dense_to_csc(A)
    for v in ("wsp", "awsp", "tcsr"):
        with S.Plan.from_dense(v, A) as pd, S.Plan.from_csc(v, 1024, 512, ptr, idx, val) as pc:
            assert pd.run_host(x).tobytes() == pc.run_host(x).tobytes(), v
            assert pd.info()["nnz"] == pc.info()["nnz"] == int(np.count_nonzero(A))


def powerlaw_csc(M, N, seed, mean8=8, cap=None):
    """BASELINE config 4 generator: column length = min(floor(8 (1-u)^(-1/2)), cap) (Pareto
    alpha=2), row ids uniform without replacement, sorted."""
    rng = np.random.default_rng(seed)
    cap = cap or M
    ln = np.minimum(np.floor(mean8 * (1.0 - rng.random(N)) ** -0.5), cap).astype(np.int64)
    ptr = np.zeros(N + 1, np.int64)
    np.cumsum(ln, out=ptr[1:])
    idx = np.empty(ptr[-1], np.int32)
    for i in range(N):
        k = ln[i]
        r = rng.integers(0, M, size=k)
        r = np.unique(r)
        while r.size < k:
            r = np.unique(np.concatenate([r, rng.integers(0, M, size=k - r.size)]))
        idx[ptr[i]:ptr[i + 1]] = r
    val = rng.uniform(-1, 1, ptr[-1]).astype(np.float32)
    val[val == 0] = 0.5
    return ptr, idx, val


def test_config3_powerlaw_scaled(S):
    """Config 4 (power-law row lengths) at a size the oracle finishes in seconds."""
    M = N = 65536 + 4096          # > 65535: 32-bit row ids, x gathered from L2
    ptr, idx, val = powerlaw_csc(M, N, 42, cap=8192)
    x = ob.gen_vector(M, 0.0, 43)
    y_ref = ob.csc_gemv(N, ptr, idx, val, x)
    s = ob.csc_gemv(N, ptr, idx, np.abs(val), np.abs(x)).astype(np.float64)
    with S.Plan.from_csc("wsp", M, N, ptr, idx, val) as p:
        y = p.run_host(x)
        assert p.info()["kernels_per_run"] == 1, "all length bins of a skewed matrix run in one merged launch"
    err = np.abs(y.astype(np.float64) - y_ref)
    assert float(np.max(err / (s + 1e-30))) <= 1e-5
    # 16-bit index variant of the same generator
    M2 = N2 = 8192
    ptr, idx, val = powerlaw_csc(M2, N2, 44, cap=2048)
    x = ob.gen_vector(M2, 0.5, 45)
    y_ref = ob.csc_gemv(N2, ptr, idx, val, x)
    s = ob.csc_gemv(N2, ptr, idx, np.abs(val), np.abs(x)).astype(np.float64)
    for v in ("wsp", "awsp", "tcsr"):
        with S.Plan.from_csc(v, M2, N2, ptr, idx, val) as p:
            err = np.abs(p.run_host(x).astype(np.float64) - y_ref)
            assert float(np.max(err / (s + 1e-30))) <= 1e-5, v


def test_config5_sharded_slab_scaled(S):
    """Config 5 shape family (very sparse, wide): one GPU's column slab at reduced size, built
    directly in sparse form; slab width auto-selects 16-bit column ids."""
    M, N, dens = 8192, 16384, 0.01
    rng = np.random.default_rng(55)
    ln = rng.binomial(M, dens, N).astype(np.int64)
    ptr = np.zeros(N + 1, np.int64)
    np.cumsum(ln, out=ptr[1:])
    idx = np.empty(ptr[-1], np.int32)
    for i in range(N):
        idx[ptr[i]:ptr[i + 1]] = np.sort(rng.choice(M, ln[i], replace=False))
    val = rng.uniform(-1, 1, ptr[-1]).astype(np.float32)
    val[val == 0] = 0.25
    x = ob.gen_vector(M, 0.5, 56)
    y_ref = ob.csc_gemv(N, ptr, idx, val, x)
    s = ob.csc_gemv(N, ptr, idx, np.abs(val), np.abs(x)).astype(np.float64)
    for v in ("awsp", "tcsr", "wsp"):
        for opts in ({}, dict(chunk_mode=1), dict(chunk_mode=2, row_splits=2), dict(chunk_mode=2, warps_per_col=4)):
            if v == "wsp" and opts:
                continue
            with S.Plan.from_csc(v, M, N, ptr, idx, val, **opts) as p:
                if v != "wsp":
                    assert p.info()["slab_cols"] > 256
                y = p.run_host(x)
                err = np.abs(y.astype(np.float64) - y_ref)
                assert float(np.max(err / (s + 1e-30))) <= 1e-5, (v, opts)
                assert p.run_host(x).tobytes() == y.tobytes(), "non-deterministic"


@pytest.mark.parametrize("M,sx", [(1, 0.0), (31, 0.5), (4096, 0.5), (14336, 0.9), (32768, 0.5),
                                  (32769, 0.5), (100000, 0.9), (1 << 20, 0.5)])
def test_compaction_bit_exact(S, M, sx):
    import torch
    x = ob.gen_vector(M, sx, 99)
    x[M // 2] = -0.0
    idx_ref, val_ref = ob.compact_x(x)
    idx, val = S.compact_x(torch.from_numpy(x).cuda())
    assert idx.cpu().numpy().tobytes() == idx_ref.tobytes()
    assert val.cpu().numpy().tobytes() == val_ref.tobytes()


def test_device_pointer_run_and_clone(S):
    import torch
    A = ob.gen_matrix(1024, 512, 0.7, 31)
    x = ob.gen_vector(1024, 0.5, 32)
    dx = torch.from_numpy(x).cuda()
    for v in VARIANTS:
        with S.Plan.from_dense(v, A) as p, p.clone() as q:
            y1 = torch.empty(512, device="cuda")
            y2 = torch.empty(512, device="cuda")
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                p.run(dx, y1)
                q.run(dx, y2)
            st.synchronize()
            assert y1.cpu().numpy().tobytes() == y2.cpu().numpy().tobytes() == p.run_host(x).tobytes()


def test_traffic_accounting(S):
    A = ob.gen_matrix(1024, 512, 0.7, 41)
    x = ob.gen_vector(1024, 0.5, 42)
    nnz = int(np.count_nonzero(A))
    mnz = int(np.count_nonzero(x))
    nnz_t = int(np.count_nonzero(A[x != 0]))
    vec = 4 * 1024 + 4 * 512
    with S.Plan.from_dense("wsp", A) as p:
        alg, phys, t = p.traffic(x)
        assert t == nnz and alg == 8 * nnz + 4 * 513 + vec and phys < alg
    with S.Plan.from_dense("asp", A) as p:
        alg, phys, t = p.traffic(x)
        assert alg == 4 * mnz * 512 + vec and phys >= alg
    for v in ("awsp", "tcsr"):
        with S.Plan.from_dense(v, A) as p:
            alg, phys, t = p.traffic(x)
            assert t == nnz_t and alg == 8 * nnz_t + 4 * 513 + vec


def test_dropin_harness_binary(S):
    """build/sparse_sgemv = test/main.cpp + the drop-in tester over the C-ABI (reference
    test/main.cpp:1-7, tester.cpp:15-34): the reference's eight launchers (cublasSgemv comparator included), the
    two csr launchers it declares but never runs, and the multi-GPU awsp launcher, on 4096x4096; every
    output inside the reference's own abs-1e-3 gate."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "sparse_sgemv")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", root, "harness"])
    env = dict(os.environ, SPMV_SEED="1234", SPMV_STRICT="1")
    out = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "========== OK ===========" in out.stdout
    assert out.stdout.count(" took ") == 11 and out.stdout.count("start to launch") == 11
    assert "cublasSgemv" in out.stdout, "the dense comparator is the library call the reference makes (cublas.cu:33)"
    assert out.stderr.strip() == ""


def test_smoke_entry(S):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as g
    g.smoke()


def test_run_scatter_into_several_buffers(S):
    """spmv_run_scatter on one GPU: the slice lands at `offset` in every destination buffer
    (what the column-sharded path does with its peers' buffers)."""
    import torch
    A = ob.gen_matrix(512, 1024, 0.7, 61)
    x = ob.gen_vector(512, 0.5, 62)
    dx = torch.from_numpy(x).cuda()
    for v in VARIANTS:
        with S.Plan.from_dense(v, A[:, 256:768]) as p:          # a 512-column slab of a 1024-column y
            ref = p.run_host(x)
            bufs = [torch.full((1024,), -7.0, device="cuda") for _ in range(3)]
            p.run_scatter(dx, [b.data_ptr() for b in bufs], 256)
            torch.cuda.synchronize()
            for b in bufs:
                h = b.cpu().numpy()
                assert h[256:768].tobytes() == ref.tobytes(), v
                assert np.all(h[:256] == -7.0) and np.all(h[768:] == -7.0), v
    with S.Plan.from_dense("wsp", A) as p:
        with pytest.raises(S.SpmvError):
            p.run_scatter(dx, [bufs[0].data_ptr()], 2)          # offset not a multiple of 4


def test_wsp_row_panels_tall(S):
    """Tall matrix (x does not fit shared memory): wsp runs per 12288-row panel and adds the panel
    sums in order; plus the 32-bit fallback when lists are short."""
    for M, N, dens, seed in ((70000, 256, 0.01, 71), (40000, 96, 0.0005, 72)):
        rng = np.random.default_rng(seed)
        from spmv_test_b200 import synth
        cp, ri, va = synth.bernoulli_csc(M, N, dens, seed)
        x = ob.gen_vector(M, 0.5, seed + 1)
        y_ref = ob.csc_gemv(N, cp, ri, va, x)
        s = ob.csc_gemv(N, cp, ri, np.abs(va), np.abs(x)).astype(np.float64)
        with S.Plan.from_csc("wsp", M, N, cp, ri, va) as p:
            y = p.run_host(x)
            assert p.run_host(x).tobytes() == y.tobytes()
            err = np.abs(y.astype(np.float64) - y_ref)
            assert float(np.max(err / (s + 1e-30))) <= 1e-5, (M, N)


def test_plan_save_load_roundtrip(S, tmp_path):
    """A plan file holds the packed format; a loaded plan computes bit-identical results and a
    damaged file is rejected."""
    A = ob.gen_matrix(768, 1024, 0.8, 81)
    x = ob.gen_vector(768, 0.5, 82)
    for v in VARIANTS:
        path = tmp_path / f"{v}.plan"
        with S.Plan.from_dense(v, A) as p:
            y = p.run_host(x)
            p.save(path)
            info = p.info()
        with S.Plan.load(path) as q:
            assert q.run_host(x).tobytes() == y.tobytes(), v
            qi = q.info()
            assert (qi["M"], qi["N"], qi["nnz"], qi["device_bytes"]) == (info["M"], info["N"], info["nnz"], info["device_bytes"])
            assert q.traffic(x) == S.Plan.load(path).traffic(x)
        raw = bytearray(path.read_bytes())
        raw[len(raw) // 2] ^= 0xFF
        bad = tmp_path / "bad.plan"
        bad.write_bytes(bytes(raw[: len(raw) - 9]))
        with pytest.raises(S.SpmvError):
            S.Plan.load(bad)
    with pytest.raises(S.SpmvError):
        S.Plan.load(tmp_path / "missing.plan")
    # tall wsp (row panels) and a very sparse awsp (multi-row chunks) survive the round trip too
    from spmv_test_b200 import synth
    cp, ri, va = synth.bernoulli_csc(40000, 256, 0.01, 83)
    x2 = ob.gen_vector(40000, 0.5, 84)
    for v in ("wsp", "awsp"):
        with S.Plan.from_csc(v, 40000, 256, cp, ri, va) as p:
            y = p.run_host(x2)
            p.save(tmp_path / "t.plan")
        with S.Plan.load(tmp_path / "t.plan") as q:
            assert q.run_host(x2).tobytes() == y.tobytes(), v


@pytest.mark.parametrize("batch", [1, 2, 3, 4, 7])
def test_run_batch_against_the_oracle_and_single_vector_calls(S, batch):
    """spmv_run_batch: every Y[b] is within the parity gate of the ORACLE and bit-identical to spmv_run on X[b]
    (wsp / asp reuse A's bytes for groups of 4 / 2 vectors, awsp / tcsr for pairs; one vector of the batch is
    all zero, one row is active in only one vector of a pair)."""
    import torch
    A = ob.gen_matrix(1024, 768, 0.7, 91)
    X = np.stack([ob.gen_vector(1024, 0.5, 100 + b) for b in range(batch)])
    if batch >= 2:
        X[1, :] = 0.0                                      # a pair with one inactive vector
        X[0, 5] = 0.25
    if batch >= 4:
        X[2, 7], X[3, 7] = 0.5, 0.0                        # a row active in one vector of the pair only
    refs_b = [(ob.sgemv_dense(A, X[b]),) + ob.sgemv_dense_f64(A, X[b]) for b in range(batch)]
    dX = torch.from_numpy(X).cuda()
    cases = [(v, {}) for v in VARIANTS] + [("awsp", {"slab_cols": 256}), ("tcsr", {"slab_cols": 256}), ("awsp", {"chunk_mode": 2}),
                                           ("awsp", {"chunk_mode": 4})]
    for v, kw in cases:
        with S.Plan.from_dense(v, A, **kw) as p:
            dY = torch.full((batch, 768), -1.0, device="cuda")
            p.run_batch(dX, dY)
            torch.cuda.synchronize()
            Y = dY.cpu().numpy()
            pairs = v in ("awsp", "tcsr") and kw.get("chunk_mode", 0) in (0, 1)   # the pair kernel: gate + reproducible
            for b in range(batch):
                y32, y64, s = refs_b[b]
                check_y(Y[b], y32, y64, s, f"batch {v} {kw} vector {b}")
                if not pairs or b == batch - 1 and batch % 2:
                    assert Y[b].tobytes() == p.run_host(X[b]).tobytes(), (v, kw, b)
            dY2 = torch.full((batch, 768), -2.0, device="cuda")
            p.run_batch(dX, dY2)
            torch.cuda.synchronize()
            assert dY2.cpu().numpy().tobytes() == Y.tobytes(), f"{v} {kw}: two batched runs differ"
    # strided X / Y (leading dimensions larger than M / N)
    for v in ("wsp", "awsp"):
        with S.Plan.from_dense(v, A) as p:
            dXp = torch.zeros((batch, 1024 + 32), device="cuda")
            dXp[:, :1024] = dX
            dYp = torch.zeros((batch, 768 + 64), device="cuda")
            p.run_batch(dXp[:, :1024], dYp[:, :768])
            torch.cuda.synchronize()
            for b in range(batch):
                y32, y64, s = refs_b[b]
                check_y(dYp[b, :768].cpu().numpy(), y32, y64, s, f"strided batch {v} vector {b}")
                if v == "wsp":
                    assert dYp[b, :768].cpu().numpy().tobytes() == p.run_host(X[b]).tobytes()
            assert float(dYp[:, 768:].abs().max()) == 0.0


def test_run_batch_config2_pairs_reuse_the_matrix(S):
    """Config 2 shape: the awsp / tcsr pair kernel against the oracle (several CTAs per slab: the partial rows
    and tickets of both vectors), reproducible run to run."""
    import torch
    A = ob.gen_matrix(4096, 14336, 0.7, 1234)
    X = np.stack([ob.gen_vector(4096, 0.5, 4321), ob.gen_vector(4096, 0.5, 4322)])
    dX = torch.from_numpy(X).cuda()
    for v in ("awsp", "tcsr"):
        with S.Plan.from_dense(v, A) as p:
            dY = torch.zeros((2, 14336), device="cuda")
            p.run_batch(dX, dY)
            torch.cuda.synchronize()
            Y = dY.cpu().numpy()
            for b in range(2):
                y32 = ob.sgemv_dense(A, X[b]); y64, s = ob.sgemv_dense_f64(A, X[b])
                check_y(Y[b], y32, y64, s, f"config 2 batched {v} vector {b}")
            dY2 = torch.zeros((2, 14336), device="cuda")
            p.run_batch(dX, dY2)
            torch.cuda.synchronize()
            assert dY2.cpu().numpy().tobytes() == Y.tobytes()


# ---- device packers (csrc/pack_dev.cu) ----------------------------------------------
DEVICE_PACK_CASES = [
    # M, N, keep-density of A, options
    (512, 512, 0.5, {}),
    (300, 1024, 0.1, {}),
    (1000, 2048, 0.3, {}),
    (64, 4096, 0.9, {}),
    (2048, 256, 0.02, {}),
    (777, 1536, 0.6, {"slab_cols": 512}),
    (640, 4096, 0.7, {"slab_cols": 4096}),
    (1, 32, 1.0, {}),
    (33, 64, 0.0, {}),
    (4096, 4096, 0.1, {}),
]


def _plan_file(plan, path):
    plan.save(path)
    return path.read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("case", DEVICE_PACK_CASES, ids=lambda c: f"{c[0]}x{c[1]}@{c[2]}")
def test_device_packers_match_host_packers(S, tmp_path, case):
    """The GPU packers (count / scan / fill / deal kernels) build the same bytes as the host
    packers: plan files are compared byte for byte, and y bit for bit."""
    import torch
    M, N, keep, opts = case
    rng = np.random.default_rng(M * 131 + N)
    A = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    A[rng.random((M, N)) >= keep] = 0.0
    if M > 8 and keep > 0:
        A[3, 5] = np.float32("nan"); A[4, 6] = -0.0; A[5, :] = 0.0      # NaN is kept, -0.0 is a zero (a14)
    x = ob.gen_vector(M, 0.5, 7)
    dA = torch.from_numpy(A).cuda()
    for v in ("wsp", "awsp", "tcsr", "asp"):
        o = {k: w for k, w in opts.items() if v in ("awsp", "tcsr")}
        with S.Plan.from_dense(v, A, pack_mode="host", **o) as ph, \
             S.Plan.from_dense(v, A, pack_mode="device", **o) as pd, \
             S.Plan.from_dense_device(v, dA, **o) as pdd:
            fh = _plan_file(ph, tmp_path / "h.plan")
            assert _plan_file(pd, tmp_path / "d.plan") == fh, (v, "host A, device packer")
            assert _plan_file(pdd, tmp_path / "dd.plan") == fh, (v, "device A")
            ih, idd = ph.info(), pdd.info()
            assert ih == idd, v
            assert ph.traffic(x) == pdd.traffic(x), v
            yh, yd = ph.run_host(x), pdd.run_host(x)
            assert np.array_equal(yh, yd, equal_nan=True), v


@pytest.mark.gpu
def test_device_packers_options_and_views(S, tmp_path):
    import torch
    # 32-bit row ids, a column-slab view with lda > N, and a tall wsp with row panels
    A = ob.gen_matrix(1024, 2048, 0.7, 91)
    dA = torch.from_numpy(A).cuda()
    with S.Plan.from_dense("wsp", A, index_bits=32, pack_mode="host") as ph, S.Plan.from_dense_device("wsp", dA, index_bits=32) as pd:
        assert _plan_file(ph, tmp_path / "a") == _plan_file(pd, tmp_path / "b")
    for v in ("wsp", "awsp", "tcsr", "asp"):
        with S.Plan.from_dense(v, A[:, 512:1280], pack_mode="host") as ph, S.Plan.from_dense_device(v, dA[:, 512:1280]) as pd:
            assert _plan_file(ph, tmp_path / "a") == _plan_file(pd, tmp_path / "b"), v
    T = ob.gen_matrix(30000, 64, 0.5, 92)
    x = ob.gen_vector(30000, 0.5, 93)
    with S.Plan.from_dense("wsp", T, pack_mode="host") as ph, S.Plan.from_dense_device("wsp", torch.from_numpy(T).cuda()) as pd:
        assert _plan_file(ph, tmp_path / "a") == _plan_file(pd, tmp_path / "b")
        assert ph.run_host(x).tobytes() == pd.run_host(x).tobytes()
    # a host pointer is rejected, as is a bad pack mode
    with pytest.raises(S.SpmvError):
        from spmv_test_b200._cabi import lib, check, VARIANTS as V
        import ctypes as C
        h = C.c_void_p()
        check(lib().spmv_plan_create_dense_device(V["awsp"], 1024, 2048, C.c_void_p(A.ctypes.data), 2048, None, C.byref(h)))
    with pytest.raises(S.SpmvError):
        S.Plan.from_dense("awsp", A, pack_mode=7)


# ---- lane-owned blocks (chunk_mode 3) --------------------------------------------------------------
LOB_SHAPES = [
    # M, N, weight sparsity, activation sparsity, slab_cols (0 = default 2048)
    (32, 32, 0.5, 0.5, 0), (1, 64, 0.0, 0.0, 1024), (1000, 512, 0.7, 0.5, 1024), (2049, 2048 + 96, 0.98, 0.5, 0),
    (5000, 4096, 0.99, 0.5, 1024), (3000, 8192, 0.995, 0.0, 4096), (70000, 512, 0.99, 0.9, 0), (4096, 4096, 0.5, 0.5, 1024),
]


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,sa,sx,W", LOB_SHAPES)
def test_lane_owned_blocks_parity(S, M, N, sa, sx, W, tmp_path):
    """chunk_mode 3: x is a multiplier, every chunk retires in one pass; same parity bar, and the
    CSR(A^T) route, a saved plan and a clone give the same bits."""
    A = ob.gen_matrix(M, N, sa, 1234)
    if M > 40:
        A[:, 3] = 0.25                                                   # one long lane stream
        A[37, :] = 0.0
    x = ob.gen_vector(M, sx, 4321)
    kw = {"slab_cols": W} if W else {}
    ys = run_all(S, A, x, variants=("awsp", "tcsr"), chunk_mode=3, **kw)
    assert ys["awsp"].tobytes() == ys["tcsr"].tobytes()
    from scipy import sparse
    c = sparse.csc_matrix(A)
    with S.Plan.from_csc("awsp", M, N, c.indptr.astype(np.int64), c.indices.astype(np.int32), c.data.astype(np.float32),
                         chunk_mode=3, **kw) as p:
        assert p.run_host(x).tobytes() == ys["awsp"].tobytes()
        p.save(tmp_path / "lob.plan")
        with p.clone() as q:
            assert q.run_host(x).tobytes() == ys["awsp"].tobytes()
        alg, phys, touched = p.traffic(x)
        assert touched == int(np.count_nonzero(A[x != 0.0])) and phys > 0
    with S.Plan.load(tmp_path / "lob.plan") as q:
        assert q.run_host(x).tobytes() == ys["awsp"].tobytes()
    raw = bytearray((tmp_path / "lob.plan").read_bytes())
    if len(raw) > 4000:
        raw[len(raw) // 2] ^= 0xFF
        (tmp_path / "bad.plan").write_bytes(bytes(raw[:-9]))
        with pytest.raises(S.SpmvError):
            S.Plan.load(tmp_path / "bad.plan")


@pytest.mark.gpu
def test_lane_owned_blocks_many_ctas_per_slab_and_forced_splits(S):
    A = ob.gen_matrix(40000, 2048, 0.99, 77)
    x = ob.gen_vector(40000, 0.5, 78)
    for o in (dict(), dict(row_splits=1), dict(row_splits=7), dict(warps_per_col=2)):
        run_all(S, A, x, variants=("awsp",), chunk_mode=3, **o)


# ---- fused activation and the decode FFN chain ---------------------------------------
@pytest.mark.gpu
def test_fused_relu_and_ffn_chain(S):
    """y = relu(x·A) through the store epilogue is bit-identical to a separate ReLU, for every
    variant; the up-projection -> ReLU -> down-projection chain (configs 2 -> 3, scaled) runs as two
    launches and matches the oracle."""
    import torch
    M, H = 512, 2048
    A1 = ob.gen_matrix(M, H, 0.7, 11)
    A2 = ob.gen_matrix(H, M, 0.7, 12)
    x = ob.gen_vector(M, 0.5, 13)
    dx = torch.from_numpy(x).cuda()
    for v in VARIANTS + ["awsp3"]:
        kw = {"chunk_mode": 3} if v == "awsp3" else {}
        with S.Plan.from_dense(v[:4] if v == "awsp3" else v, A1, **kw) as p:
            y = torch.empty(H, device="cuda"); yr = torch.empty(H, device="cuda")
            p.run(dx, y); p.run(dx, yr, act="relu")
            torch.cuda.synchronize()
            yn = y.cpu().numpy()
            assert np.array_equal(yr.cpu().numpy(), np.where(yn < 0, np.float32(0), yn)), v
            assert np.any(yn < 0)
            y32a, y64a, sa = refs(A1, x)                                 # ... and the rectified output against the ORACLE's y
            check_y(yr.cpu().numpy(), np.where(y32a < 0, np.float32(0), y32a), np.where(y32a < 0, 0.0, y64a), sa, f"relu {v}")
    with S.Plan.from_dense("wsp", A1) as up, S.Plan.from_dense("awsp", A2) as down:
        h = torch.empty(H, device="cuda"); z = torch.empty(M, device="cuda")
        up.run(dx, h, act="relu")
        down.run(h, z)                                                  # its prologue compacts the rectified h
        torch.cuda.synchronize()
        hn = h.cpu().numpy()
        assert 0.3 < float(np.mean(hn == 0)) < 0.7                      # about half of the intermediate is inactive
        y32, y64, s = refs(A2, hn)
        check_y(z.cpu().numpy(), y32, y64, s, "ffn chain")
    with pytest.raises(S.SpmvError):
        with S.Plan.from_dense("asp", A1) as p:
            p.run(dx, torch.empty(H, device="cuda"), act=7)


# ---- row strips (chunk_mode 4) ---------------------------------------------------------------------
STRIP_SHAPES = [
    # M, N, weight sparsity, activation sparsity, strip columns (0 = from the density)
    (32, 32, 0.5, 0.5, 0), (1, 64, 0.0, 0.0, 32), (1000, 512, 0.7, 0.5, 32), (2049, 2048 + 96, 0.98, 0.5, 0),
    (5000, 4096, 0.99, 0.5, 0), (3000, 8192, 0.995, 0.0, 2112), (70000, 512, 0.99, 0.9, 0), (4100, 4096, 0.9, 0.5, 128),
    (300, 40000, 0.99, 0.3, 0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,sa,sx,sw", STRIP_SHAPES)
def test_row_strips_parity(S, M, N, sa, sx, sw, tmp_path):
    """chunk_mode 4: one row segment per 32-lane window, rows with x == 0 never read; same parity bar,
    and the CSR(A^T) route, a saved plan, a clone and a scattered run give the same bits."""
    import torch
    A = ob.gen_matrix(M, N, sa, 1234)
    if M > 40:
        A[:, 3] = 0.25                                                   # a column every row touches
        A[37, :] = 0.0                                                   # an empty row
        A[38, :] = -0.5                                                  # a full row: segments longer than a window
    x = ob.gen_vector(M, sx, 4321)
    if M > 40:
        x[38] = 0.75
    kw = {"slab_cols": sw} if sw else {}
    ys = run_all(S, A, x, variants=("awsp",), chunk_mode=4, **kw)
    from scipy import sparse
    c = sparse.csc_matrix(A)
    with S.Plan.from_csc("awsp", M, N, c.indptr.astype(np.int64), c.indices.astype(np.int32), c.data.astype(np.float32),
                         chunk_mode=4, **kw) as p:
        assert p.run_host(x).tobytes() == ys["awsp"].tobytes()
        p.save(tmp_path / "strips.plan")
        with p.clone() as q:
            assert q.run_host(x).tobytes() == ys["awsp"].tobytes()
        alg, phys, touched = p.traffic(x)
        assert touched == int(np.count_nonzero(A[x != 0.0])) and phys > 0
        # the scatter epilogue (multi-GPU path) through the same kernels
        bufs = [torch.full((N + 64,), -7.0, device="cuda") for _ in range(2)]
        p.run_scatter(torch.from_numpy(x).cuda(), [b.data_ptr() for b in bufs], 32)
        torch.cuda.synchronize()
        for b in bufs:
            assert b[32:32 + N].cpu().numpy().tobytes() == ys["awsp"].tobytes()
            assert float(b[:32].max()) == -7.0 and float(b[32 + N:].max()) == -7.0
        yr = torch.empty(N, device="cuda")
        p.run(torch.from_numpy(x).cuda(), yr, act="relu")
        torch.cuda.synchronize()
        assert np.array_equal(yr.cpu().numpy(), np.where(ys["awsp"] < 0, np.float32(0), ys["awsp"]))
    with S.Plan.load(tmp_path / "strips.plan") as q:
        assert q.run_host(x).tobytes() == ys["awsp"].tobytes()
    raw = bytearray((tmp_path / "strips.plan").read_bytes())
    if len(raw) > 4000:
        raw[len(raw) // 2] ^= 0xFF
        (tmp_path / "bad.plan").write_bytes(bytes(raw[:-9]))
        with pytest.raises(S.SpmvError):
            S.Plan.load(tmp_path / "bad.plan")


@pytest.mark.gpu
def test_row_strips_row_ranges_and_zero_x(S):
    """Forced CTAs per band (one: direct stores of y; many: the partial rows + the reduce kernel; more
    rows than one compaction pass per CTA), an all-zero x, and x with a single active row."""
    A = ob.gen_matrix(9000, 4096, 0.99, 77)
    x = ob.gen_vector(9000, 0.5, 78)
    outs = []
    for o in (dict(), dict(row_splits=1), dict(row_splits=2), dict(row_splits=7), dict(row_splits=140)):
        outs.append(run_all(S, A, x, variants=("awsp",), chunk_mode=4, **o)["awsp"])
    x0 = np.zeros(9000, np.float32)
    run_all(S, A, x0, variants=("awsp",), chunk_mode=4)
    x1 = x0.copy(); x1[8999] = 2.0
    y = run_all(S, A, x1, variants=("awsp",), chunk_mode=4)["awsp"]
    assert np.array_equal(y, 2.0 * A[8999])


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,density,sw", [(4096, 8192, 0.01, 0), (1000, 2048 + 96, 0.03, 64), (70000, 512, 0.01, 0),
                                            (300, 40000, 0.01, 0), (33, 64, 0.5, 32), (5, 32, 0.0, 0)])
def test_row_strips_device_packer_matches_host_packer(S, tmp_path, M, N, density, sw):
    """spmv_plan_create_csc_device: CSR(A^T) arrays in device memory packed by kernels; the plan file is the
    host packer's byte for byte, and bad input (unsorted / repeated / out-of-range rows) is rejected."""
    import torch
    from spmv_test_b200 import synth
    if density > 0:
        cp, ri, va = synth.bernoulli_csc(M, N, density, seed=77)
    else:
        cp, ri, va = np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32)
    if va.size > 10:
        va[3] = 0.0                                        # an explicit zero is not stored
        va[7] = -0.0
    kw = {"slab_cols": sw} if sw else {}
    x = ob.gen_vector(M, 0.5, 78)
    with S.Plan.from_csc("awsp", M, N, cp, ri, va, chunk_mode=4, **kw) as ph, \
            S.Plan.from_csc_device("awsp", M, N, torch.from_numpy(cp).cuda(), torch.from_numpy(ri).cuda(), torch.from_numpy(va).cuda(),
                                   chunk_mode=4, **kw) as pd:
        ph.save(tmp_path / "host.plan")
        pd.save(tmp_path / "dev.plan")
        assert (tmp_path / "host.plan").read_bytes() == (tmp_path / "dev.plan").read_bytes()
        assert ph.run_host(x).tobytes() == pd.run_host(x).tobytes()
        assert ph.traffic(x) == pd.traffic(x) and ph.info()["nnz"] == pd.info()["nnz"]
    if ri.size > 4:
        for bad in ("swap", "repeat", "range"):
            r2 = ri.copy()
            k = int(cp[np.argmax(np.diff(cp) >= 2)])       # first column with two entries
            if bad == "swap":
                r2[k], r2[k + 1] = r2[k + 1], r2[k]
            elif bad == "repeat":
                r2[k + 1] = r2[k]
            else:
                r2[k] = M
            with pytest.raises(S.SpmvError):
                S.Plan.from_csc_device("awsp", M, N, torch.from_numpy(cp).cuda(), torch.from_numpy(r2).cuda(),
                                       torch.from_numpy(va).cuda(), chunk_mode=4, **kw)
    with pytest.raises(S.SpmvError):                       # other forms are packed on the host
        S.Plan.from_csc_device("awsp", M, N, torch.from_numpy(cp).cuda(), torch.from_numpy(ri).cuda(), torch.from_numpy(va).cuda())


# ---- BASELINE configs 4 and 5 at FULL size, oracle on sampled columns ------------------------------
def _sampled_oracle(N, cp, ri, va, x, n_cols, seed):
    """y of `n_cols` seeded columns through the oracle (fp32 sequential csr_naive.cu:14-22 semantics + fp64)."""
    rng = np.random.default_rng(seed)
    cols = np.sort(rng.choice(N, size=n_cols, replace=False))
    lens = (cp[cols + 1] - cp[cols]).astype(np.int64)
    scp = np.zeros(n_cols + 1, np.int64)
    np.cumsum(lens, out=scp[1:])
    take = np.concatenate([np.arange(cp[c], cp[c + 1]) for c in cols])
    sri, sva = ri[take], va[take]
    y32 = ob.csc_gemv(n_cols, scp, sri, sva, x)
    xa = x.astype(np.float64)[sri] * sva.astype(np.float64)
    seg = np.repeat(np.arange(n_cols), lens)
    return cols, y32, np.bincount(seg, weights=xa, minlength=n_cols), np.bincount(seg, weights=np.abs(xa), minlength=n_cols)


@pytest.mark.gpu
def test_config5_full_size_slab_against_sampled_oracle(S):
    """One GPU's slab of BASELINE config 5 exactly as bench.py builds it (65536 x 131072, 1 %, seed 5000, x seed 4321):
    the host packer, the device packer and the automatic choice of the row-strip form give the same bits, 4096
    sampled columns are inside the parity gate, two runs are bit-identical."""
    import torch
    from spmv_test_b200 import synth
    M, N = 65536, 131072
    cp, ri, va = synth.bernoulli_csc(M, N, 0.01, seed=5000)
    x = synth.gen_vector(M, 0.5, seed=4321)
    cols, y32, y64, s = _sampled_oracle(N, cp, ri, va, x, 4096, 900)
    with S.Plan.from_csc_device("awsp", M, N, torch.from_numpy(cp).cuda(), torch.from_numpy(ri).cuda(), torch.from_numpy(va).cuda(),
                                chunk_mode=4) as pd:
        y = pd.run_host(x)
        check_y(y[cols], y32, y64, s, "config 5 slab, row strips (device packer)")
        assert pd.run_host(x).tobytes() == y.tobytes()
        info = pd.info()
        assert info["slab_cols"] == 2048 and info["kernels_per_run"] == 2 and info["nnz"] == int(np.count_nonzero(va))
        alg, phys, touched = pd.traffic(x)
        assert touched == int(np.count_nonzero(x[ri] != 0)) and alg == 8.0 * touched + 4.0 * (N + 1) + 4.0 * M + 4.0 * N
    with S.Plan.from_csc("awsp", M, N, cp, ri, va) as pa:                   # chunk_mode auto -> row strips on this density
        assert pa.info()["slab_cols"] == 2048 and pa.info()["index_bits"] == 32
        assert pa.run_host(x).tobytes() == y.tobytes()


@pytest.mark.gpu
def test_config4_full_size_against_sampled_oracle(S):
    """BASELINE config 4 exactly as bench.py builds it (1M x 1M power-law columns, seed 42, dense x): 4096 sampled
    columns inside the parity gate, one merged launch, two runs bit-identical."""
    from spmv_test_b200 import synth
    M = N = 1 << 20
    cp, ri, va = synth.powerlaw_csc(M, N, seed=42)
    x = synth.gen_vector(M, 0.0, seed=7)
    cols, y32, y64, s = _sampled_oracle(N, cp, ri, va, x, 4096, 404)
    with S.Plan.from_csc("wsp", M, N, cp, ri, va) as p:
        y = p.run_host(x)
        check_y(y[cols], y32, y64, s, "config 4 wsp")
        assert p.info()["kernels_per_run"] == 1 and p.run_host(x).tobytes() == y.tobytes()

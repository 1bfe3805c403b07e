"""Pins the CPU oracle (oracle/spmv_oracle.c) and the product's host re-implementation of the
reference layouts (spmv_ref_pack) to the REFERENCE:
  * against tests/golden/ref_golden.json + ref_golden_y.npz, minted from the reference's own
    host code compiled in place (tests/golden/make_golden.py) — travels with the repo;
  * live against oracle/_ref/libspmv_ref_cpu.so when it exists (authoring container).
Everything here is bit-exact: integer/bitmap/index streams, packed values, and the fp32
sequential product (same order of operations, -ffp-contract=off)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_bindings as ob
from golden.make_golden import CASES, LAYOUTS, case_inputs, digest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ref_golden.json")) as f:
    GOLD = json.load(f)["cases"]
GOLD_Y = np.load(os.path.join(HERE, "golden", "ref_golden_y.npz"))


def packed_digests(p):
    return {"i32_a": digest(p.i32_a) if p.i32_a is not None and p.i32_a.size else None,
            "i32_b": digest(p.i32_b) if p.i32_b is not None and p.i32_b.size else None,
            "u32": digest(p.u32) if p.u32 is not None and p.u32.size else None,
            "f32": digest(p.f32) if p.f32 is not None and p.f32.size else None,
            "aux": [int(v) for v in p.aux]}


def norm(d):
    """golden digests of empty/absent streams are equivalent"""
    return {k: (None if (v is None or (isinstance(v, str) and v.endswith(":0"))) else v) for k, v in d.items()}


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_inputs_reproduce(case):
    name, M, N, sa, sx, seed = case
    A, x = case_inputs(M, N, sa, sx, seed)
    assert digest(A) == GOLD[name]["A"] and digest(x) == GOLD[name]["x"]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_sgemv_matches_reference_bits(case):
    name, M, N, sa, sx, seed = case
    A, x = case_inputs(M, N, sa, sx, seed)
    y = ob.sgemv_dense(A, x)
    assert y.tobytes() == GOLD_Y[name].tobytes()
    assert digest(y) == GOLD[name]["y"]


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_packers_match_reference_bits(case, layout):
    name, M, N, sa, sx, seed = case
    A, _ = case_inputs(M, N, sa, sx, seed)
    assert norm(packed_digests(ob.pack(layout, A))) == norm(GOLD[name]["layouts"][layout])


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_product_ref_layouts_match_reference_bits(case, layout):
    """spmv_ref_pack (what the drop-in CSRMatrix/TCSRMatrix/... classes hold) vs the reference."""
    import spmv_test_b200 as S
    name, M, N, sa, sx, seed = case
    A, _ = case_inputs(M, N, sa, sx, seed)
    assert norm(packed_digests(S.ref_pack(layout, A))) == norm(GOLD[name]["layouts"][layout])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_decodes_equal_dense_loop(case):
    """Every reference kernel's index math, restated on the CPU with plain ascending-row
    accumulation (gpu_order = 0), reproduces SgemvCPU bit for bit."""
    name, M, N, sa, sx, seed = case
    A, x = case_inputs(M, N, sa, sx, seed)
    y = GOLD_Y[name]
    for layout in ("csr", "tcsr", "asp", "awsp", "awsp_ref"):
        if layout in ("asp", "awsp", "awsp_ref") and M % 128:
            continue        # the reference kernels' own precondition
        assert ob.decode_gemv(layout, A, x).tobytes() == y.tobytes(), layout
    if M % 1024 == 0:       # wsp kernel needs M % 1024 == 0 (wsp.cu:23-26)
        assert ob.decode_gemv("wsp", A, x).tobytes() == y.tobytes()


def test_oracle_gpu_order_emulations_within_tolerance():
    """gpu_order = 1 reproduces the reference kernels' own association (lane partials, shuffle
    tree, 4-warp sum, fmaf); different order, same tolerance gate."""
    from parity import check_y
    A = ob.gen_matrix(1024, 256, 0.7, 77)
    x = ob.gen_vector(1024, 0.5, 78)
    y32 = ob.sgemv_dense(A, x)
    y64, s = ob.sgemv_dense_f64(A, x)
    for layout, ver in (("csr", 0), ("tcsr", 0), ("wsp", 0), ("asp", 0), ("asp", 2), ("awsp", 0), ("awsp", 2), ("awsp_ref", 0)):
        check_y(ob.decode_gemv(layout, A, x, gpu_order=1, version=ver), y32, y64, s, f"{layout} v{ver}")


def test_compaction_oracle_semantics():
    x = np.array([0.0, -0.0, 1.0, np.nan, -2.0, 0.0, np.inf, 1e-45], np.float32)
    idx, val = ob.compact_x(x)
    assert idx.tolist() == [2, 3, 4, 6, 7]      # -0.0 dropped, NaN / Inf / denormal kept (asp.cu:23)
    assert val.tobytes() == x[[2, 3, 4, 6, 7]].tobytes()


@pytest.mark.skipif(not ob.have_ref_cpu(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("M,N,sa", [(64, 64, 0.5), (4096, 128, 0.9), (512, 1024, 0.7), (128, 2048, 0.99)])
def test_live_against_compiled_reference(M, N, sa):
    import spmv_test_b200 as S
    A = ob.gen_matrix(M, N, sa, M + N)
    x = ob.gen_vector(M, 0.5, M * 3 + 1)
    assert ob.sgemv_dense(A, x).tobytes() == ob.ref_sgemv_cpu(A, x).tobytes()
    for layout in LAYOUTS:
        ref = ob.ref_pack(layout, A)
        ok, f = ob.packed_equal(ob.pack(layout, A), ref)
        assert ok, f"oracle {layout}: {f}"
        ok, f = ob.packed_equal(S.ref_pack(layout, A), ref)
        assert ok, f"product {layout}: {f}"

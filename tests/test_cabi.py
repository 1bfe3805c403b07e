"""The C-ABI library loads, exports every symbol include/spmv_b200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "spmv_b200.h")).read()
    return sorted(set(re.findall(r"SPMV_API[^;(]*?\b(spmv_\w+)\s*\(", src)))


def test_header_symbols_exported():
    import spmv_test_b200 as S
    from spmv_test_b200 import _cabi
    syms = header_symbols()
    assert len(syms) >= 19
    assert sorted(_cabi.SYMBOLS) == syms, "binding list and header disagree"
    L = S.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} not exported by libspmv_b200.so"
    out = subprocess.check_output(["nm", "-D", "--defined-only", S.LIB_PATH], text=True)
    exported = set(re.findall(r" T (\w+)", out))
    assert set(syms) <= exported
    assert all(e.startswith("spmv_") for e in exported), f"unexpected exports: {sorted(exported - set(syms))[:5]}"


def test_library_does_not_link_the_oracle():
    import spmv_test_b200 as S
    out = subprocess.check_output(["ldd", S.LIB_PATH], text=True)
    assert "oracle" not in out and "spmv_ref" not in out
    strings = subprocess.check_output(["strings", S.LIB_PATH], text=True)
    assert "orc_sgemv_dense" not in strings and "liboracle" not in strings


def test_python_host_side_never_imports_oracle():
    pkg = os.path.join(ROOT, "spmv_test_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn


def test_abi_version_and_errors():
    import spmv_test_b200 as S
    L = S.lib()
    assert L.spmv_abi_version() == 1
    h = C.c_void_p()
    A = np.zeros((32, 48), np.float32)
    rc = L.spmv_plan_create_dense(99, 32, 32, C.c_void_p(A.ctypes.data), 32, None, C.byref(h))
    assert rc == -1 and b"variant" in L.spmv_last_error()
    rc = L.spmv_plan_create_dense(0, 32, 48, C.c_void_p(A.ctypes.data), 48, None, C.byref(h))
    assert rc == -2 and b"multiple of 32" in L.spmv_last_error()
    assert L.spmv_run(None, None, None, None) == -1
    assert L.spmv_plan_info(None, None) == -1


def test_no_cpu_fallback_without_gpu():
    import spmv_test_b200 as S
    if S.lib().spmv_device_count() > 0:
        pytest.skip("a GPU is present")
    A = np.ones((64, 64), np.float32)
    for v in ("wsp", "asp", "awsp", "tcsr"):
        with pytest.raises(S.SpmvError) as e:
            S.Plan.from_dense(v, A)
        assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_partition_columns():
    import spmv_test_b200 as S
    b = S.column_bounds(1048576, 8, 4096)
    assert b.tolist() == [g * 131072 for g in range(9)]
    b = S.column_bounds(14336, 4, 256)
    assert b[0] == 0 and b[-1] == 14336 and all(x % 256 == 0 for x in b) and all(np.diff(b) > 0)
    b = S.column_bounds(96, 8, 32)                      # more GPUs than slabs: empty tails allowed
    assert b[-1] == 96 and all(np.diff(b) >= 0)
    # nnz-balanced: heavy first half
    col_ptr = np.zeros(1025, np.int64)
    col_ptr[1:] = np.cumsum(np.where(np.arange(1024) < 512, 30, 10))
    b = S.column_bounds(1024, 2, 32, col_ptr)
    assert b[1] < 512 and b[1] % 32 == 0
    assert abs(col_ptr[b[1]] - col_ptr[-1] / 2) <= 30 * 32


def test_harness_binary_links_dropin_headers():
    exe = os.path.join(ROOT, "build", "sparse_sgemv")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "harness"])
    out = subprocess.check_output(["nm", "-C", exe], text=True)
    for sym in ("wsp_gemv_gpu(int, int, float*, float*, float*, int)", "awsp_ref_gemv_gpu(int, int, float*, float*, float*)",
                "csr_tiling_gemv_gpu(int, int, float*, float*, float*)", "SparseSgemvTester::RunTest()",
                "CSRMatrix::PrintCSR()"):
        assert sym in out, sym

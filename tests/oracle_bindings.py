"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference
itself compiled in place (oracle/_ref/libspmv_ref_{cpu,gpu}.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_CPU_SO = os.path.join(ROOT, "oracle", "_ref", "libspmv_ref_cpu.so")
REF_GPU_SO = os.path.join(ROOT, "oracle", "_ref", "libspmv_ref_gpu.so")

LAYOUTS = {"csr": 0, "tcsr": 1, "wsp": 2, "asp": 3, "awsp": 4, "awsp_ref": 5}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------------
# oracle
# ------------------------------------------------------------------------------------------
_orc = None


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = C.CDLL(ORACLE_SO)
        L.orc_sgemv_dense.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.orc_sgemv_dense_f64.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f64p, _f64p]
        L.orc_compact_x.argtypes = [C.c_int, _f32p, _i32p, _f32p]
        L.orc_compact_x.restype = C.c_int
        L.orc_count_nnz.argtypes = [C.c_int64, _f32p]
        L.orc_count_nnz.restype = C.c_int64
        L.orc_pack_csr.argtypes = [C.c_int, C.c_int, _f32p, _i32p, _i32p, _f32p]
        L.orc_pack_csr.restype = C.c_int
        L.orc_pack_tcsr.argtypes = [C.c_int, C.c_int, _f32p, _i32p, _u32p, _f32p]
        L.orc_pack_tcsr.restype = C.c_int
        L.orc_pack_wsp.argtypes = [C.c_int, C.c_int, _f32p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_pack_wsp.restype = C.c_int
        L.orc_pack_asp.argtypes = [C.c_int, C.c_int, _f32p, _f32p]
        L.orc_pack_awsp.argtypes = [C.c_int, C.c_int, _f32p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_pack_awsp.restype = C.c_int
        L.orc_pack_awsp_ref.argtypes = [C.c_int, C.c_int, _f32p, C.c_void_p, C.c_void_p, _i32p]
        L.orc_csr_naive_gemv.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_int]
        L.orc_csr_tiling_gemv.argtypes = [C.c_int, C.c_int, _f32p, _i32p, _u32p, _f32p, _f32p, C.c_int]
        L.orc_wsp_gemv.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _f32p, _f32p, _f32p, C.c_int]
        L.orc_asp_gemv.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.orc_awsp_gemv.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.orc_awsp_ref_gemv.argtypes = [C.c_int, C.c_int, _u32p, _f32p, _i32p, _f32p, _f32p, C.c_int]
        L.orc_csc_gemv.argtypes = [C.c_int64, _i64p, _i32p, _f32p, _f32p, _f32p, C.c_int]
        L.orc_dense_to_csc.argtypes = [C.c_int, C.c_int, _f32p, _i64p, C.c_void_p, C.c_void_p]
        L.orc_dense_to_csc.restype = C.c_int64
        _orc = L
    return _orc


def sgemv_dense(A, x):
    """tester.cpp:36-45 restated (oracle/spmv_oracle.c: orc_sgemv_dense)."""
    A = _f32(A); x = _f32(x)
    M, N = A.shape
    y = np.empty(N, np.float32)
    oracle().orc_sgemv_dense(M, N, A, x, y)
    return y


def sgemv_dense_f64(A, x):
    A = _f32(A); x = _f32(x)
    M, N = A.shape
    y = np.zeros(N, np.float64)
    s = np.zeros(N, np.float64)
    oracle().orc_sgemv_dense_f64(M, N, A, x, y, s)
    return y, s


def compact_x(x):
    x = _f32(x)
    idx = np.empty(max(x.size, 1), np.int32)
    val = np.empty(max(x.size, 1), np.float32)
    n = oracle().orc_compact_x(x.size, x, idx, val)
    return idx[:n].copy(), val[:n].copy()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pack(layout, A):
    """Oracle packers.  Returns a namespace with the same fields as spmv_ref_packed_t."""
    A = _f32(A)
    M, N = A.shape
    L = oracle()
    out = SimpleNamespace(i32_a=None, i32_b=None, u32=None, f32=None, aux=[0, 0, 0, 0])
    nnz = int(L.orc_count_nnz(A.size, A.reshape(-1)))
    if layout == "csr":
        out.i32_a = np.empty(N, np.int32)
        out.i32_b = np.empty(max(nnz, 1), np.int32)
        out.f32 = np.empty(max(nnz, 1), np.float32)
        n = L.orc_pack_csr(M, N, A, out.i32_a, out.i32_b, out.f32)
        assert n == nnz
        out.i32_b = out.i32_b[:nnz]; out.f32 = out.f32[:nnz]
    elif layout == "tcsr":
        out.i32_a = np.empty((M // 32) * (N // 32) + 1, np.int32)
        out.u32 = np.empty(max(M * N // 32, 1), np.uint32)
        out.f32 = np.empty(max(nnz, 1), np.float32)
        n = L.orc_pack_tcsr(M, N, A, out.i32_a, out.u32, out.f32)
        assert n == nnz
        out.u32 = out.u32[: M * N // 32]; out.f32 = out.f32[:nnz]
    elif layout == "wsp":
        nzm = L.orc_pack_wsp(M, N, A, None, None, 0)
        out.u32 = np.empty(max(M * N // 32, 1), np.uint32)
        out.f32 = np.empty(max(N * nzm, 1), np.float32)
        L.orc_pack_wsp(M, N, A, _ptr(out.u32), _ptr(out.f32), nzm)
        out.u32 = out.u32[: M * N // 32]; out.f32 = out.f32[: N * nzm]
        out.aux[0] = nzm; out.aux[1] = N
    elif layout == "asp":
        out.f32 = np.empty(M * N, np.float32)
        L.orc_pack_asp(M, N, A, out.f32)
    elif layout == "awsp":
        bk = L.orc_pack_awsp(M, N, A, None, None, 0)
        out.u32 = np.empty(max(M * N // 32, 1), np.uint32)
        out.f32 = np.empty(max(M * N // 1024 * bk, 1), np.float32)
        L.orc_pack_awsp(M, N, A, _ptr(out.u32), _ptr(out.f32), bk)
        out.u32 = out.u32[: M * N // 32]; out.f32 = out.f32[: M * N // 1024 * bk]
        out.aux[0] = bk
    elif layout == "awsp_ref":
        off = np.zeros(4, np.int32)
        L.orc_pack_awsp_ref(M, N, A, None, None, off)
        out.i32_a = off
        out.u32 = np.empty(max(M * N // 32, 1), np.uint32)
        out.f32 = np.empty(max(N // 32 * int(off[3]), 1), np.float32)
        L.orc_pack_awsp_ref(M, N, A, _ptr(out.u32), _ptr(out.f32), off)
        out.u32 = out.u32[: M * N // 32]; out.f32 = out.f32[: N // 32 * int(off[3])]
    else:
        raise ValueError(layout)
    return out


def decode_gemv(layout, A, x, gpu_order=0, version=0):
    """Pack with the oracle packer, then run the oracle's restatement of the matching
    reference kernel's decode."""
    A = _f32(A); x = _f32(x)
    M, N = A.shape
    L = oracle()
    p = pack(layout, A)
    y = np.empty(N, np.float32)
    if layout == "csr":
        L.orc_csr_naive_gemv(M, N, p.f32 if p.f32.size else np.zeros(1, np.float32), p.f32.size,
                             p.i32_b if p.i32_b.size else np.zeros(1, np.int32), p.i32_a, x, y, gpu_order)
    elif layout == "tcsr":
        L.orc_csr_tiling_gemv(M, N, p.f32 if p.f32.size else np.zeros(1, np.float32), p.i32_a, p.u32, x, y, gpu_order)
    elif layout == "wsp":
        L.orc_wsp_gemv(M, N, p.aux[0], p.u32, p.f32 if p.f32.size else np.zeros(1, np.float32), x, y, gpu_order)
    elif layout == "asp":
        L.orc_asp_gemv(M, N, p.f32, x, y, version, gpu_order)
    elif layout == "awsp":
        L.orc_awsp_gemv(M, N, p.aux[0], p.u32, p.f32 if p.f32.size else np.zeros(1, np.float32), x, y, version, gpu_order)
    elif layout == "awsp_ref":
        L.orc_awsp_ref_gemv(M, N, p.u32, p.f32 if p.f32.size else np.zeros(1, np.float32), p.i32_a, x, y, gpu_order)
    else:
        raise ValueError(layout)
    return y


def dense_to_csc(A):
    A = _f32(A)
    M, N = A.shape
    L = oracle()
    ptr = np.empty(N + 1, np.int64)
    nnz = int(L.orc_dense_to_csc(M, N, A, ptr, None, None))
    idx = np.empty(max(nnz, 1), np.int32)
    val = np.empty(max(nnz, 1), np.float32)
    L.orc_dense_to_csc(M, N, A, ptr, _ptr(idx), _ptr(val))
    return ptr, idx[:nnz], val[:nnz]


def csc_gemv(N, col_ptr, row_idx, vals, x, threads=1):
    y = np.empty(N, np.float32)
    oracle().orc_csc_gemv(N, np.ascontiguousarray(col_ptr, np.int64),
                          np.ascontiguousarray(row_idx, np.int32) if len(row_idx) else np.zeros(1, np.int32),
                          _f32(vals) if len(vals) else np.zeros(1, np.float32), _f32(x), y, threads)
    return y


# ------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref), when it was built
# ------------------------------------------------------------------------------------------
class _RefPacked(C.Structure):
    _fields_ = [("i32_a", C.POINTER(C.c_int32)), ("n_i32_a", C.c_int64),
                ("i32_b", C.POINTER(C.c_int32)), ("n_i32_b", C.c_int64),
                ("u32", C.POINTER(C.c_uint32)), ("n_u32", C.c_int64),
                ("f32", C.POINTER(C.c_float)), ("n_f32", C.c_int64),
                ("aux", C.c_int32 * 4)]


def packed_struct_to_ns(s):
    def grab(p, n, dt):
        if not p or n == 0:
            return np.zeros(0, dt) if p else None
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)
    return SimpleNamespace(i32_a=grab(s.i32_a, s.n_i32_a, np.int32), i32_b=grab(s.i32_b, s.n_i32_b, np.int32),
                           u32=grab(s.u32, s.n_u32, np.uint32), f32=grab(s.f32, s.n_f32, np.float32),
                           aux=list(s.aux))


_ref_cpu = None
_ref_gpu = None


def have_ref_cpu():
    return os.path.exists(REF_CPU_SO)


def have_ref_gpu():
    return os.path.exists(REF_GPU_SO)


def ref_cpu():
    global _ref_cpu
    if _ref_cpu is None:
        L = C.CDLL(REF_CPU_SO)
        L.ref_sgemv_cpu.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.ref_pack.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, C.POINTER(_RefPacked)]
        L.ref_pack.restype = C.c_int
        L.ref_packed_free.argtypes = [C.POINTER(_RefPacked)]
        _ref_cpu = L
    return _ref_cpu


def ref_sgemv_cpu(A, x):
    """The reference's own SgemvCPU (tester.cpp:36-45), compiled from /root/reference."""
    A = _f32(A); x = _f32(x)
    M, N = A.shape
    y = np.empty(N, np.float32)
    ref_cpu().ref_sgemv_cpu(M, N, A, x, y)
    return y


def ref_pack(layout, A):
    A = _f32(A)
    M, N = A.shape
    s = _RefPacked()
    rc = ref_cpu().ref_pack(LAYOUTS[layout], M, N, A, C.byref(s))
    assert rc == 0
    ns = packed_struct_to_ns(s)
    ref_cpu().ref_packed_free(C.byref(s))
    return ns


REF_GPU_KERNELS = {"cublas": 0, "wsp": 1, "asp": 2, "awsp": 3, "awsp_ref": 4, "csr_naive": 5,
                   "csr_tiling": 6, "naive": 7, "tiling": 8}


def ref_gpu_gemv(kernel, A, x, version=0):
    """Run one of the reference's own GPU launchers (recompiled for sm_100a).  Returns
    (y, milliseconds printed by the reference's TIME_KERNEL)."""
    global _ref_gpu
    if _ref_gpu is None:
        L = C.CDLL(REF_GPU_SO)
        L.ref_gpu_gemv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.ref_gpu_gemv.restype = C.c_float
        _ref_gpu = L
    A = _f32(A); x = _f32(x)
    M, N = A.shape
    y = np.zeros(N, np.float32)
    ms = _ref_gpu.ref_gpu_gemv(REF_GPU_KERNELS[kernel], version, M, N, A, x, y)
    return y, float(ms)


# ------------------------------------------------------------------------------------------
# seeded synthetic inputs (the reference's generator, tester.cpp:103-121,151-167, made
# explicit: keep an element iff U[0,1) > sparsity, value U(-1,1); numpy's PCG64 instead of
# an unseeded mt19937)
# ------------------------------------------------------------------------------------------
def gen_matrix(M, N, sparsity, seed):
    rng = np.random.default_rng(seed)
    keep = rng.random((M, N)) > sparsity
    vals = rng.uniform(-1.0, 1.0, (M, N)).astype(np.float32)
    return np.where(keep, vals, np.float32(0.0)).astype(np.float32)


def gen_vector(M, sparsity, seed):
    rng = np.random.default_rng(seed)
    keep = rng.random(M) > sparsity
    vals = rng.uniform(-1.0, 1.0, M).astype(np.float32)
    return np.where(keep, vals, np.float32(0.0)).astype(np.float32)


def packed_equal(a, b):
    for f in ("i32_a", "i32_b", "u32", "f32"):
        x, y = getattr(a, f), getattr(b, f)
        if (x is None or x.size == 0) and (y is None or y.size == 0):
            continue
        if x is None or y is None or x.shape != y.shape or x.tobytes() != y.tobytes():
            return False, f
    if list(a.aux) != list(b.aux):
        return False, "aux"
    return True, ""

"""Plan objects over the C-ABI (include/spmv_b200.h).

A plan is the pack-once / run-many form of one reference launcher
(`*_gemv_gpu(M, N, A_host, X_host, Y_host[, version])`, reference src/include/kernel.hpp:8-17,
which re-packs and re-uploads A on every call).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from ._cabi import LAYOUTS, VARIANTS, Options, PackedDump, PlanInfo, RefPacked, check, lib


PACK_MODES = {"auto": 0, "host": 1, "device": 2}


def _opts(**kw):
    if not kw:
        return None
    if isinstance(kw.get("pack_mode"), str):
        kw["pack_mode"] = PACK_MODES[kw["pack_mode"]]
    o = Options()
    o.struct_size = C.sizeof(Options)
    for k, v in kw.items():
        if v is not None:
            setattr(o, k, int(v))
    return C.byref(o)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


class Plan:
    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    # ---- construction ----------------------------------------------------------------------
    @classmethod
    def from_dense(cls, variant, A, **opts):
        """A: 2-D float32 array, row-major; a column slab view `A[:, a:b]` of a C-contiguous
        matrix is accepted as is (leading dimension = the parent's row length)."""
        A = np.asarray(A)
        if A.dtype != np.float32 or A.ndim != 2:
            raise TypeError("A must be a 2-D float32 array")
        M, N = A.shape
        if A.size and (A.strides[1] != 4 or A.strides[0] % 4 or A.strides[0] < 4 * N):
            A = np.ascontiguousarray(A)
        lda = A.strides[0] // 4 if A.size and M > 1 else max(N, 1)
        if lda < N:
            lda = N
        h = C.c_void_p()
        check(lib().spmv_plan_create_dense(VARIANTS[variant], M, N, _ptr(A), lda, _opts(**opts), C.byref(h)))
        p = cls(h.value)
        p._keep = None
        return p

    @classmethod
    def from_dense_device(cls, variant, A, **opts):
        """A: 2-D float32 CUDA tensor (row-major, unit column stride) on the current device: the
        formats are built by the device packers (csrc/pack_dev.cu), bit-identical to from_dense."""
        if A.dim() != 2 or not A.is_cuda or str(A.dtype) != "torch.float32":
            raise TypeError("A must be a 2-D float32 CUDA tensor")
        M, N = A.shape
        if A.numel() and (A.stride(1) != 1 or A.stride(0) < N):
            A = A.contiguous()
        lda = A.stride(0) if A.numel() and M > 1 else max(N, 1)
        h = C.c_void_p()
        check(lib().spmv_plan_create_dense_device(VARIANTS[variant], M, N, C.c_void_p(A.data_ptr() if A.numel() else 0),
                                                  max(lda, N), _opts(**opts), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_csc(cls, variant, M, N, col_ptr, row_idx, values, **opts):
        """CSR(A^T) input: per output column i the entries (row, value), rows ascending —
        the orientation of the reference's CSRMatrix (matrix_csr.cpp:8-22), 64-bit pointers
        with the N+1 sentinel."""
        col_ptr = np.ascontiguousarray(col_ptr, np.int64)
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        values = np.ascontiguousarray(values, np.float32)
        if col_ptr.size != N + 1:
            raise ValueError("col_ptr needs N+1 entries")
        h = C.c_void_p()
        check(lib().spmv_plan_create_csc(VARIANTS[variant], M, N, _ptr(col_ptr), _ptr(row_idx), _ptr(values),
                                         _opts(**opts), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_csc_device(cls, variant, M, N, col_ptr, row_idx, values, **opts):
        """CSR(A^T) input as CUDA tensors (int64 col_ptr[N+1], int32 row_idx, float32 values) on the
        current device: packed by kernels (row-strip form: variant "awsp", chunk_mode=4), bit-identical
        to from_csc."""
        import torch
        if not (col_ptr.is_cuda and row_idx.is_cuda and values.is_cuda):
            raise TypeError("col_ptr / row_idx / values must be CUDA tensors")
        if col_ptr.dtype != torch.int64 or row_idx.dtype != torch.int32 or values.dtype != torch.float32:
            raise TypeError("col_ptr int64, row_idx int32, values float32")
        if col_ptr.numel() != N + 1:
            raise ValueError("col_ptr needs N+1 entries")
        col_ptr, row_idx, values = col_ptr.contiguous(), row_idx.contiguous(), values.contiguous()
        h = C.c_void_p()
        check(lib().spmv_plan_create_csc_device(VARIANTS[variant], M, N, C.c_void_p(col_ptr.data_ptr()),
                                                C.c_void_p(row_idx.data_ptr() if row_idx.numel() else 0),
                                                C.c_void_p(values.data_ptr() if values.numel() else 0), _opts(**opts), C.byref(h)))
        return cls(h.value)

    def save(self, path):
        """Writes the packed format to `path` (no re-packing on load)."""
        check(lib().spmv_plan_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path, **opts):
        h = C.c_void_p()
        check(lib().spmv_plan_load(str(path).encode(), _opts(**opts), C.byref(h)))
        return cls(h.value)

    def clone(self):
        h = C.c_void_p()
        check(lib().spmv_plan_clone(self._h, C.byref(h)))
        return Plan(h.value)

    def close(self):
        if self._h:
            lib().spmv_plan_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- queries ---------------------------------------------------------------------------
    def info(self):
        i = PlanInfo()
        check(lib().spmv_plan_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in PlanInfo._fields_ if k != "reserved"}

    def traffic(self, x):
        """(algorithmic bytes, physical bytes, non-zeros touched) of one call with this x."""
        x = np.ascontiguousarray(x, np.float32)
        a, p, n = C.c_double(), C.c_double(), C.c_int64()
        check(lib().spmv_plan_traffic(self._h, _ptr(x), C.byref(a), C.byref(p), C.byref(n)))
        return a.value, p.value, n.value

    # ---- execution -------------------------------------------------------------------------
    def run(self, d_x, d_y, stream=None, act=None):
        """Asynchronous y = x·A on device tensors (torch CUDA tensors or raw pointers);
        act="relu" fuses the activation into the stores of y (spmv_run_act)."""
        px = d_x if isinstance(d_x, int) else d_x.data_ptr()
        py = d_y if isinstance(d_y, int) else d_y.data_ptr()
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        if act is None:
            check(lib().spmv_run(self._h, C.c_void_p(px), C.c_void_p(py), C.c_void_p(stream)))
        else:
            code = {"none": 0, "relu": 1}[act] if isinstance(act, str) else int(act)
            check(lib().spmv_run_act(self._h, C.c_void_p(px), C.c_void_p(py), code, C.c_void_p(stream)))

    def run_batch(self, d_X, d_Y, stream=None):
        """Y[b] = X[b]·A for a batch of activation vectors (2-D CUDA tensors, row-major)."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        check(lib().spmv_run_batch(self._h, int(d_X.shape[0]), C.c_void_p(d_X.data_ptr()), int(d_X.stride(0)),
                                   C.c_void_p(d_Y.data_ptr()), int(d_Y.stride(0)), C.c_void_p(stream)))

    def run_scatter(self, d_x, dst_ptrs, offset, multicast_ptr=0, stream=None):
        """y slice of this rank -> [offset, offset+N) of every buffer in dst_ptrs (device pointers
        of all ranks' full-y buffers), or through the multicast alias when given: the all-gather
        fused into the kernel epilogue."""
        px = d_x if isinstance(d_x, int) else d_x.data_ptr()
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        arr = (C.c_void_p * len(dst_ptrs))(*[int(q) for q in dst_ptrs])
        check(lib().spmv_run_scatter(self._h, C.c_void_p(px), len(dst_ptrs), arr, C.c_void_p(int(multicast_ptr) or None),
                                     int(offset), C.c_void_p(stream)))

    def run_host(self, x, y=None, timing=False):
        """Host-buffer call (H2D x, kernels, D2H y, synchronise) — the per-call part of a
        reference launcher.  Returns y (and the kernels' device milliseconds if timing)."""
        info = self.info()
        x = np.ascontiguousarray(x, np.float32)
        if x.size != info["M"]:
            raise ValueError("x has the wrong length")
        if y is None:
            y = np.empty(info["N"], np.float32)
        ms = C.c_float(-1.0)
        check(lib().spmv_run_host(self._h, _ptr(x), _ptr(y), C.byref(ms) if timing else None))
        return (y, ms.value) if timing else y

    def run_host_ptr(self, x_ptr, y_ptr):
        """Same with raw host pointers (pinned buffers owned by the caller)."""
        check(lib().spmv_run_host(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), None))


def compact_x(d_x, stream=None):
    """Device activation compaction: returns (idx int32[count], val float32[count]) tensors."""
    import torch
    M = d_x.numel()
    idx = torch.empty(max(M, 1), dtype=torch.int32, device=d_x.device)
    val = torch.empty(max(M, 1), dtype=torch.float32, device=d_x.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=d_x.device)
    nb = lib().spmv_compact_x_scratch_bytes(M)
    scratch = torch.empty(max(nb, 4), dtype=torch.uint8, device=d_x.device)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    check(lib().spmv_compact_x(C.c_void_p(d_x.data_ptr()), M, C.c_void_p(idx.data_ptr()), C.c_void_p(val.data_ptr()),
                               C.c_void_p(cnt.data_ptr()), C.c_void_p(scratch.data_ptr()), nb, C.c_void_p(stream)))
    n = int(cnt.item())
    return idx[:n], val[:n]


def ref_pack(layout, A):
    """The reference's host layouts (bit-exact re-implementation, CPU only): what the drop-in
    CSRMatrix / TCSRMatrix / WSPMatrix / ASPMatrix / AWSPMatrix / AWSPRefMatrix classes hold."""
    A = np.ascontiguousarray(A, np.float32)
    M, N = A.shape
    s = RefPacked()
    check(lib().spmv_ref_pack(LAYOUTS[layout], M, N, _ptr(A), C.byref(s)))

    def grab(p, n, dt):
        if not p:
            return None
        if n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)

    out = SimpleNamespace(i32_a=grab(s.i32_a, s.n_i32_a, np.int32), i32_b=grab(s.i32_b, s.n_i32_b, np.int32),
                          u32=grab(s.u32, s.n_u32, np.uint32), f32=grab(s.f32, s.n_f32, np.float32),
                          aux=list(s.aux))
    lib().spmv_ref_packed_free(C.byref(s))
    return out


def pack_dump(variant, A=None, csc=None, shape=None, **opts):
    """Host image of the device format (no GPU needed): for the CPU tests of the packers.
    Pass a dense `A`, or `csc=(col_ptr, row_idx, values)` with `shape=(M, N)`."""
    d = PackedDump()
    if A is not None:
        A = np.ascontiguousarray(A, np.float32)
        M, N = A.shape
        check(lib().spmv_pack_dump_dense(VARIANTS[variant], M, N, _ptr(A), max(N, 1), _opts(**opts), C.byref(d)))
    else:
        M, N = shape
        cp = np.ascontiguousarray(csc[0], np.int64)
        ri = np.ascontiguousarray(csc[1], np.int32)
        va = np.ascontiguousarray(csc[2], np.float32)
        check(lib().spmv_pack_dump_csc(VARIANTS[variant], M, N, _ptr(cp), _ptr(ri), _ptr(va), _opts(**opts), C.byref(d)))

    def arr(p, n, dt):
        if not p or n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)

    idt = {8: np.uint8, 16: np.uint16, 32: np.uint32}[d.index_bits]
    nidx = d.idx_bytes // np.dtype(idt).itemsize
    idx = np.zeros(0, idt)
    if d.idx and nidx:
        idx = np.frombuffer(C.string_at(d.idx, d.idx_bytes), dtype=idt).copy()
    out = SimpleNamespace(variant=variant, index_bits=d.index_bits, slab_cols=d.slab_cols, slabs=d.slabs,
                          row_blocks=d.row_blocks, block_rows=d.block_rows, M=d.M, N=d.N, nnz=d.nnz, groups=d.groups,
                          vals=arr(d.vals, d.n_vals, np.float32), idx=idx,
                          off=arr(d.off, d.n_off, np.uint32), rel=arr(d.rel, d.n_rel, np.uint16))
    lib().spmv_pack_dump_free(C.byref(d))
    return out

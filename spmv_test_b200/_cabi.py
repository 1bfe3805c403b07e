"""ctypes binding of include/spmv_b200.h.  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SPMV_B200_LIB: development override (tools/trace_build.sh builds an instrumented library)
LIB_PATH = os.environ.get("SPMV_B200_LIB") or os.path.join(HERE, "lib", "libspmv_b200.so")

VARIANTS = {"wsp": 0, "asp": 1, "awsp": 2, "tcsr": 3}
LAYOUTS = {"csr": 0, "tcsr": 1, "wsp": 2, "asp": 3, "awsp": 4, "awsp_ref": 5}

# every symbol include/spmv_b200.h declares (tests/test_cabi.py checks the header against this)
SYMBOLS = [
    "spmv_abi_version", "spmv_last_error", "spmv_device_count", "spmv_set_device", "spmv_plan_create_dense", "spmv_plan_create_dense_device",
    "spmv_plan_create_csc", "spmv_plan_create_csc_device", "spmv_plan_info", "spmv_plan_destroy", "spmv_plan_clone", "spmv_plan_save", "spmv_plan_load",
    "spmv_plan_traffic", "spmv_run", "spmv_run_act", "spmv_run_batch", "spmv_run_scatter", "spmv_run_host", "spmv_compact_x",
    "spmv_compact_x_scratch_bytes", "spmv_partition_columns", "spmv_ref_pack",
    "spmv_ref_packed_free", "spmv_pack_dump_dense", "spmv_pack_dump_csc", "spmv_pack_dump_free",
    "spmv_mg_block_bytes", "spmv_mg_create", "spmv_mg_destroy", "spmv_mg_ipc_handle", "spmv_mg_connect_ipc",
    "spmv_mg_connect_ptrs", "spmv_mg_add_plan", "spmv_mg_run", "spmv_mg_run_host", "spmv_mg_status",
    "spmv_mg_create_group", "spmv_mg_group_run_host",
]


class SpmvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libspmv_b200 error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("row_splits", C.c_int32), ("warps_per_col", C.c_int32),
                ("index_bits", C.c_int32), ("slab_cols", C.c_int32), ("chunk_mode", C.c_int32),
                ("pack_mode", C.c_int32), ("reserved", C.c_int32 * 1)]


class PlanInfo(C.Structure):
    _fields_ = [("variant", C.c_int32), ("M", C.c_int64), ("N", C.c_int64), ("nnz", C.c_int64),
                ("device_bytes", C.c_int64), ("scratch_bytes", C.c_int64),
                ("kernels_per_run", C.c_int32), ("grid_x", C.c_int32), ("grid_y", C.c_int32),
                ("block", C.c_int32), ("smem_bytes", C.c_int32), ("index_bits", C.c_int32),
                ("row_splits", C.c_int32), ("warps_per_col", C.c_int32), ("slab_cols", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class RefPacked(C.Structure):
    _fields_ = [("i32_a", C.POINTER(C.c_int32)), ("n_i32_a", C.c_int64),
                ("i32_b", C.POINTER(C.c_int32)), ("n_i32_b", C.c_int64),
                ("u32", C.POINTER(C.c_uint32)), ("n_u32", C.c_int64),
                ("f32", C.POINTER(C.c_float)), ("n_f32", C.c_int64),
                ("aux", C.c_int32 * 4)]


class PackedDump(C.Structure):
    _fields_ = [("variant", C.c_int32), ("index_bits", C.c_int32), ("slab_cols", C.c_int32), ("slabs", C.c_int32),
                ("row_blocks", C.c_int32), ("block_rows", C.c_int32),
                ("M", C.c_int64), ("N", C.c_int64), ("nnz", C.c_int64), ("groups", C.c_int64),
                ("vals", C.POINTER(C.c_float)), ("n_vals", C.c_int64),
                ("idx", C.c_void_p), ("idx_bytes", C.c_int64),
                ("off", C.POINTER(C.c_uint32)), ("n_off", C.c_int64),
                ("rel", C.POINTER(C.c_uint16)), ("n_rel", C.c_int64)]


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it was not built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `make lib` (or __graft_entry__.build()); "
                           "this package has no CPU or PyTorch fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.spmv_abi_version.restype = i32
    L.spmv_last_error.restype = C.c_char_p
    L.spmv_device_count.restype = i32
    L.spmv_set_device.argtypes = [i32]
    L.spmv_plan_create_dense.argtypes = [i32, i64, i64, vp, i64, C.POINTER(Options), C.POINTER(vp)]
    L.spmv_plan_create_dense_device.argtypes = [i32, i64, i64, vp, i64, C.POINTER(Options), C.POINTER(vp)]
    L.spmv_plan_create_csc.argtypes = [i32, i64, i64, vp, vp, vp, C.POINTER(Options), C.POINTER(vp)]
    L.spmv_plan_create_csc_device.argtypes = [i32, i64, i64, vp, vp, vp, C.POINTER(Options), C.POINTER(vp)]
    L.spmv_plan_info.argtypes = [vp, C.POINTER(PlanInfo)]
    L.spmv_plan_destroy.argtypes = [vp]
    L.spmv_plan_destroy.restype = None
    L.spmv_plan_clone.argtypes = [vp, C.POINTER(vp)]
    L.spmv_plan_save.argtypes = [vp, C.c_char_p]
    L.spmv_plan_load.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(vp)]
    L.spmv_plan_traffic.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.spmv_run.argtypes = [vp, vp, vp, vp]
    L.spmv_run_act.argtypes = [vp, vp, vp, i32, vp]
    L.spmv_run_batch.argtypes = [vp, i32, vp, i64, vp, i64, vp]
    L.spmv_run_scatter.argtypes = [vp, vp, i32, C.POINTER(vp), vp, i64, vp]
    L.spmv_run_host.argtypes = [vp, vp, vp, C.POINTER(C.c_float)]
    L.spmv_compact_x.argtypes = [vp, i64, vp, vp, vp, vp, C.c_size_t, vp]
    L.spmv_compact_x_scratch_bytes.argtypes = [i64]
    L.spmv_compact_x_scratch_bytes.restype = C.c_size_t
    L.spmv_partition_columns.argtypes = [i64, i32, i64, vp, vp]
    L.spmv_ref_pack.argtypes = [i32, i32, i32, vp, C.POINTER(RefPacked)]
    L.spmv_ref_packed_free.argtypes = [C.POINTER(RefPacked)]
    L.spmv_ref_packed_free.restype = None
    L.spmv_pack_dump_dense.argtypes = [i32, i64, i64, vp, i64, C.POINTER(Options), C.POINTER(PackedDump)]
    L.spmv_pack_dump_csc.argtypes = [i32, i64, i64, vp, vp, vp, C.POINTER(Options), C.POINTER(PackedDump)]
    L.spmv_pack_dump_free.argtypes = [C.POINTER(PackedDump)]
    L.spmv_pack_dump_free.restype = None
    L.spmv_mg_block_bytes.argtypes = [i64]
    L.spmv_mg_block_bytes.restype = C.c_size_t
    L.spmv_mg_create.argtypes = [i64, i64, i32, i32, vp, C.POINTER(vp)]
    L.spmv_mg_destroy.argtypes = [vp]
    L.spmv_mg_destroy.restype = None
    L.spmv_mg_ipc_handle.argtypes = [vp, vp]
    L.spmv_mg_connect_ipc.argtypes = [vp, vp]
    L.spmv_mg_connect_ptrs.argtypes = [vp, C.POINTER(vp), vp]
    L.spmv_mg_add_plan.argtypes = [vp, vp, i64]
    L.spmv_mg_run.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.spmv_mg_run_host.argtypes = [vp, vp, vp, i64, i64]
    L.spmv_mg_status.argtypes = [vp]
    L.spmv_mg_create_group.argtypes = [i64, i64, i32, C.POINTER(i32), C.POINTER(vp)]
    L.spmv_mg_group_run_host.argtypes = [C.POINTER(vp), i32, vp, vp]
    if L.spmv_abi_version() != 1:
        raise RuntimeError("libspmv_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SpmvError(rc, lib().spmv_last_error().decode(errors="replace"))

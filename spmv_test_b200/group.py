"""Column-sharded groups over the C-ABI (spmv_mg_*, include/spmv_b200.h).

A `Group` is one rank's share of a column-sharded matrix: its plans (slabs of columns at a column
offset of the full y) plus the shared block that holds two alternating copies of the full y and the
arrival flags.  `run` launches the rank's plans back to back; their epilogues store into every
rank's y and a one-warp arrival kernel replaces the all-gather's synchronisation (no host round
trip, CUDA-graph capturable).  With world == 1 it is "several slabs, one y" on one GPU.

The reference is single-GPU; this is the multi-GPU row of the scope table.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._cabi import check, lib


class Group:
    def __init__(self, M, N_total, rank=0, world=1, block_ptr=0):
        """block_ptr: device pointer of this rank's shared block (`Group.block_bytes(N_total)` bytes,
        e.g. a symmetric-memory tensor); 0 lets the library allocate it (exchange `ipc_handle()`s)."""
        self.M, self.N, self.rank, self.world = int(M), int(N_total), int(rank), int(world)
        h = C.c_void_p()
        check(lib().spmv_mg_create(self.M, self.N, self.rank, self.world, C.c_void_p(int(block_ptr) or None), C.byref(h)))
        self._h = h
        self._plans = []
        self.npad = (self.N + 63) // 64 * 64
        self.calls = 0

    @staticmethod
    def block_bytes(N_total):
        return int(lib().spmv_mg_block_bytes(int(N_total)))

    def close(self):
        if getattr(self, "_h", None):
            lib().spmv_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- wiring ------------------------------------------------------------------------------------
    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().spmv_mg_ipc_handle(self._h, buf))
        return buf.raw

    def connect_ipc(self, handles):
        """handles: the 64-byte handles of all ranks, in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 64 * self.world
        check(lib().spmv_mg_connect_ipc(self._h, C.c_char_p(blob)))

    def connect_ptrs(self, ptrs, multicast_ptr=0):
        """ptrs[r]: rank r's block as mapped into this process (symmetric memory buffer_ptrs)."""
        arr = (C.c_void_p * self.world)(*[int(q) for q in ptrs])
        check(lib().spmv_mg_connect_ptrs(self._h, arr, C.c_void_p(int(multicast_ptr) or None)))

    def add(self, plan, col_offset):
        check(lib().spmv_mg_add_plan(self._h, plan._h, int(col_offset)))
        self._plans.append(plan)                           # keep the plan alive

    # ---- execution ---------------------------------------------------------------------------------
    def run(self, d_x, stream=None):
        """Asynchronous.  Returns the device pointer of this rank's copy of the full y (N floats),
        complete when `stream` reaches this point; it stays valid until the call after the next."""
        px = d_x if isinstance(d_x, int) else d_x.data_ptr()
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        y = C.c_void_p()
        check(lib().spmv_mg_run(self._h, C.c_void_p(px), C.c_void_p(stream), C.byref(y)))
        self.calls += 1
        return int(y.value or 0)

    def run_host(self, x, y=None, y_begin=0, y_count=None):
        """H2D x, run, D2H y[y_begin : y_begin + y_count], synchronise.  x / y: numpy arrays or raw
        host pointers (ints; then y_count is required)."""
        if y_count is None:
            y_count = self.N - y_begin
        if isinstance(x, int):
            px, py = x, int(y or 0)
            out = None
        else:
            x = np.ascontiguousarray(x, np.float32)
            out = np.empty(y_count, np.float32) if y is None else y
            px, py = x.ctypes.data, out.ctypes.data
        check(lib().spmv_mg_run_host(self._h, C.c_void_p(px), C.c_void_p(py or None), int(y_begin), int(y_count)))
        self.calls += 1
        return out

    def status(self):
        check(lib().spmv_mg_status(self._h))


def local_group(M, N_total, devices):
    """Several devices of THIS process (spmv_mg_create_group): returns one Group per device."""
    n = len(devices)
    dev = (C.c_int32 * n)(*[int(d) for d in devices])
    out = (C.c_void_p * n)()
    check(lib().spmv_mg_create_group(int(M), int(N_total), n, dev, out))
    groups = []
    for i in range(n):
        g = Group.__new__(Group)
        g.M, g.N, g.rank, g.world = int(M), int(N_total), i, n
        g._h, g._plans, g.npad, g.calls = C.c_void_p(out[i]), [], (int(N_total) + 63) // 64 * 64, 0
        groups.append(g)
    return groups


def group_run_host(groups, x, y=None):
    """One call on all devices of a local group from one host thread; returns the full y."""
    x = np.ascontiguousarray(x, np.float32)
    if y is None:
        y = np.empty(groups[0].N, np.float32)
    arr = (C.c_void_p * len(groups))(*[g._h for g in groups])
    check(lib().spmv_mg_group_run_host(arr, len(groups), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data)))
    for g in groups:
        g.calls += 1
    return y

"""Column-slab partitioner for the multi-GPU config (BASELINE config 5).

The reference is single-GPU.  Outputs are independent, so A is sharded by
contiguous column slabs: rank g owns outputs bounds[g]..bounds[g+1], x is replicated (broadcast
once per call when only rank 0 holds it), every rank runs the single-GPU kernel on its slab and
the Y slices are joined with one all-gather.  No reduction crosses GPUs.

One process per GPU; `torch.distributed` is the plumbing (NCCL on GPUs, gloo in the CPU tests,
which inject the local compute since the product has no CPU path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._cabi import check, lib


def column_bounds(N, parts, align=256, col_ptr=None):
    """bounds[g]..bounds[g+1] = outputs of rank g (spmv_partition_columns in the C-ABI)."""
    b = np.zeros(parts + 1, np.int64)
    cp = None if col_ptr is None else np.ascontiguousarray(col_ptr, np.int64)
    check(lib().spmv_partition_columns(int(N), int(parts), int(align),
                                       C.c_void_p(cp.ctypes.data) if cp is not None else C.c_void_p(0),
                                       C.c_void_p(b.ctypes.data)))
    return b


class ShardedSgemv:
    """y = x·A with A column-sharded over the ranks of a process group.

    `local_run(d_x, d_y_slice)` computes this rank's slice; by default it is `plan.run`.

    join = "nccl"   the Y slices are joined by one `all_gather_into_tensor` (slices padded to the
                    widest slab).
    join = "fused"  the all-gather is fused into the kernel epilogue: y lives in symmetric memory
                    (torch.distributed._symmetric_memory: every rank maps every rank's buffer), the
                    kernel's final stores go straight to this rank's slice of EVERY rank's y — one
                    multimem.st through the NVSwitch multicast address when NVLS is available,
                    else one peer store per rank over NVLink (`spmv_run_scatter`) — and a
                    symmetric-memory barrier replaces the collective.  Two y buffers alternate so
                    a fast rank cannot overwrite a y its peer is still reading.
    """

    def __init__(self, bounds, rank, world, plan=None, local_run=None, group=None, device=None, join="nccl"):
        import torch
        self.bounds = [int(b) for b in bounds]
        self.rank, self.world, self.group = rank, world, group
        self.N = self.bounds[-1]
        self.width = max(self.bounds[g + 1] - self.bounds[g] for g in range(world))
        self.plan = plan
        self.join = join
        if local_run is None:
            if plan is None:
                raise ValueError("ShardedSgemv needs a plan (there is no CPU compute path)")
            local_run = plan.run
        self.local_run = local_run
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.equal = all(self.bounds[g + 1] - self.bounds[g] == self.width for g in range(world))
        if join == "fused":
            if plan is None or world < 2:
                raise ValueError("join='fused' needs a plan and at least two ranks")
            if any(b % 4 for b in self.bounds):
                raise ValueError("slab bounds must be multiples of 4 columns")
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            self._len = (self.N + 3) // 4 * 4
            self.y_sym = symm.empty(2 * self._len, dtype=torch.float32, device=self.device)
            self.y_sym.zero_()
            grp = group if group is not None else dist.group.WORLD
            self.hdl = symm.rendezvous(self.y_sym, grp)
            self.ptrs = [int(q) for q in self.hdl.buffer_ptrs]
            # NVSwitch multicast alias of the buffers (NVLS): 0 when the allocation has none
            mc = 0
            try:
                mc = int(self.hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
            if os.environ.get("SPMV_NO_MULTICAST"):
                mc = 0
            self.multicast = mc
            self.flip = 0
            torch.cuda.synchronize()
            self.hdl.barrier(channel=0)
        else:
            self.y_local = torch.zeros(self.width, dtype=torch.float32, device=self.device)
            self.y_all = torch.zeros(self.width * world, dtype=torch.float32, device=self.device)

    def run(self, d_x, x_on_all_ranks=True):
        """Returns the full y (length N) on every rank."""
        import torch.distributed as dist
        if not x_on_all_ranks:
            dist.broadcast(d_x, src=0, group=self.group)
        if self.join == "fused":
            base = self.flip * self._len
            self.flip ^= 1
            ptrs = [q + 4 * base for q in self.ptrs]
            mc = self.multicast + 4 * base if self.multicast else 0
            self.plan.run_scatter(d_x, ptrs, self.bounds[self.rank], mc)
            self.hdl.barrier(channel=0)                    # every rank's stores have landed
            return self.y_sym[base: base + self.N]
        self.local_run(d_x, self.y_local)
        if self.world == 1:
            return self.y_local[: self.N]
        dist.all_gather_into_tensor(self.y_all, self.y_local, group=self.group)
        if self.equal:
            return self.y_all[: self.N]
        import torch
        parts = [self.y_all[g * self.width: g * self.width + self.bounds[g + 1] - self.bounds[g]]
                 for g in range(self.world)]
        return torch.cat(parts)

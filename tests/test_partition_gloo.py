"""Multi-GPU host logic on CPU: world_size = 2 over gloo.  The product has no CPU compute
path, so the per-rank kernel is replaced by an injected callable (the oracle's CSR(A^T)
product on that rank's slab) — what is under test is the partitioner, the slab bookkeeping
and the all-gather join of ShardedSgemv."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, align, balanced, x_everywhere, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_bindings as ob
        import spmv_test_b200 as S
        M = 512
        A = ob.gen_matrix(M, N, 0.9, 100)
        A[:, : N // 2] *= (np.random.default_rng(1).random((M, N // 2)) > 0.5)   # lighter first half
        x = ob.gen_vector(M, 0.5, 101)
        ptr, idx, val = ob.dense_to_csc(A)
        bounds = S.column_bounds(N, world, align, ptr if balanced else None)
        a, b = int(bounds[rank]), int(bounds[rank + 1])
        sl_ptr = ptr[a:b + 1] - ptr[a]
        sl_idx, sl_val = idx[ptr[a]:ptr[b]], val[ptr[a]:ptr[b]]

        def local_run(d_x, d_y):          # stands in for plan.run on this rank's slab
            y = ob.csc_gemv(b - a, sl_ptr, sl_idx, sl_val, d_x.numpy())
            d_y.zero_()
            d_y[: b - a] = torch.from_numpy(y)

        sh = S.ShardedSgemv(bounds, rank, world, local_run=local_run, device=torch.device("cpu"))
        d_x = torch.from_numpy(x.copy()) if (x_everywhere or rank == 0) else torch.zeros(M)
        y = sh.run(d_x, x_on_all_ranks=x_everywhere).numpy()
        ref = ob.csc_gemv(N, ptr, idx, val, x)
        ok = y.shape == (N,) and y.tobytes() == ref.tobytes()
        q.put((rank, bool(ok), [int(v) for v in bounds]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,align,balanced,x_everywhere", [
    (1024, 256, False, True),     # equal slabs: the all-gather output is y itself
    (1280, 256, False, True),     # unequal slabs: padded gather + trim
    (1024, 32, True, False),      # nnz-balanced bounds, x broadcast from rank 0
])
def test_sharded_sgemv_world2(N, align, balanced, x_everywhere):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, align, balanced, x_everywhere, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    b = res[0][2]
    assert b[0] == 0 and b[-1] == N and all(v % align == 0 for v in b[:-1])
    if balanced:
        assert b[1] > N // 2, "nnz-balanced split should give the lighter half more columns"


def test_sharded_requires_a_compute_path():
    import spmv_test_b200 as S
    with pytest.raises(ValueError):
        S.ShardedSgemv([0, 32, 64], 0, 2, device=torch.device("cpu"))

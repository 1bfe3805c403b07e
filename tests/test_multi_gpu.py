"""Two-rank GPU test of the column-sharded path: NCCL all-gather join and the fused epilogue
(peer / multicast stores from inside the kernel + symmetric-memory barrier).  Needs two GPUs;
skipped otherwise (the CPU-side logic is covered by tests/test_partition_gloo.py)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, variant, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle_bindings as ob
        import spmv_test_b200 as S
        M, N = 1024, 2048
        A = ob.gen_matrix(M, N, 0.8, 300)
        x = ob.gen_vector(M, 0.5, 301)
        y32 = ob.sgemv_dense(A, x)
        y64, s = ob.sgemv_dense_f64(A, x)
        bounds = S.column_bounds(N, world, 256)
        a, b = int(bounds[rank]), int(bounds[rank + 1])
        plan = S.Plan.from_dense(variant, A[:, a:b])
        dx = torch.from_numpy(x).cuda()
        out = {}
        for join in ("nccl", "fused"):
            sh = S.ShardedSgemv(bounds, rank, world, plan=plan, join=join)
            ys = []
            for _ in range(3):                       # three calls: exercises the alternating buffers
                ys.append(sh.run(dx).clone())
            torch.cuda.synchronize()
            y = ys[-1].cpu().numpy()
            assert all(v.cpu().numpy().tobytes() == y.tobytes() for v in ys)
            err = np.abs(y.astype(np.float64) - y64)
            out[join] = (float(np.max(err / (s + 1e-30))), y.tobytes() == ys[0].cpu().numpy().tobytes(), y)
        same = out["nccl"][2].tobytes() == out["fused"][2].tobytes()
        q.put((rank, out["nccl"][0], out["fused"][0], same, getattr(sh, "multicast", 0) != 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["awsp", "wsp", "asp", "tcsr"])
def test_sharded_two_gpus(variant):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, variant, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, e_nccl, e_fused, same, mc in res:
        assert e_nccl <= 1e-5 and e_fused <= 1e-5, res
        assert same, "fused epilogue and NCCL join disagree"

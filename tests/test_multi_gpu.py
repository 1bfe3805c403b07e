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


# ---- spmv_mg_* groups (C-ABI) ----------------------------------------------------------------------
def test_group_several_slabs_on_one_gpu():
    """world = 1: three slabs of different variants and widths into one y — the N = 1 form of the
    column-sharded matrix (bench.py's config 5 on a single GPU); the y buffers alternate call by call,
    the host-buffer call returns any range of y."""
    import torch
    sys.path.insert(0, HERE)
    import oracle_bindings as ob
    from parity import check_y
    import spmv_test_b200 as S
    M, widths = 3000, (512, 1024, 2048 + 64)
    N = sum(widths)
    A = ob.gen_matrix(M, N, 0.97, 400)
    x = ob.gen_vector(M, 0.5, 401)
    y32 = ob.sgemv_dense(A, x)
    y64, s = ob.sgemv_dense_f64(A, x)
    block = torch.zeros(S.Group.block_bytes(N) // 4, dtype=torch.float32, device="cuda")
    with S.Group(M, N, 0, 1, block.data_ptr()) as g:
        off, plans = 0, []
        for w, (v, kw) in zip(widths, (("wsp", {}), ("awsp", {"chunk_mode": 4}), ("asp", {}))):
            p = S.Plan.from_dense(v, A[:, off:off + w], **kw)
            g.add(p, off)
            plans.append(p)
            off += w
        dx = torch.from_numpy(x).cuda()
        ptrs = []
        for _ in range(3):
            ptrs.append(g.run(dx))
        torch.cuda.synchronize()
        assert ptrs[0] == ptrs[2] != ptrs[1], "two y buffers alternate"
        ys = [block[(q - block.data_ptr()) // 4:][:N].cpu().numpy() for q in ptrs[1:]]
        assert ys[0].tobytes() == ys[1].tobytes()
        check_y(ys[0], y32, y64, s, "group of three slabs")
        # each slab alone gives the same bits
        off = 0
        for w, p in zip(widths, plans):
            assert p.run_host(x).tobytes() == ys[0][off:off + w].tobytes()
            off += w
        part = g.run_host(x, y_begin=512, y_count=1024)
        assert part.tobytes() == ys[0][512:1536].tobytes()
        # the copy-back pipeline (one event per slab, second stream): ranges that cross slab boundaries, the full range,
        # a single column, repeated calls (both y buffers)
        for b, c in ((0, N), (100, 1000), (508, 8), (1536 - 4, 2048), (N - 1, 1), (0, 1)):
            for _ in range(2):
                assert g.run_host(x, y_begin=b, y_count=c).tobytes() == ys[0][b:b + c].tobytes(), (b, c)
        g.status()
        with pytest.raises(S.SpmvError):
            g.add(plans[0], N - 4)                        # does not fit
    with S.Group(M, N) as g2:                             # library-owned block: the IPC handle is available
        assert len(g2.ipc_handle()) == 64


def test_local_group_two_devices_of_one_process():
    """spmv_mg_create_group / spmv_mg_group_run_host: what a single-process C++ caller (the reference's
    harness is one process) uses to reach several GPUs: peer access, stores into the peer's y, arrival flags."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, HERE)
    import oracle_bindings as ob
    from parity import check_y
    import spmv_test_b200 as S
    M, N = 2048, 4096
    A = ob.gen_matrix(M, N, 0.9, 410)
    x = ob.gen_vector(M, 0.5, 411)
    y32 = ob.sgemv_dense(A, x)
    y64, s = ob.sgemv_dense_f64(A, x)
    groups = S.local_group(M, N, [0, 1])
    try:
        for d, g in enumerate(groups):
            torch.cuda.set_device(d)
            for k in range(2):                            # two slabs per device
                a = (2 * d + k) * (N // 4)
                g.add(S.Plan.from_dense("awsp" if k else "wsp", A[:, a:a + N // 4]), a)
        for _ in range(3):
            y = S.group_run_host(groups, x)
        check_y(y, y32, y64, s, "two devices of one process")
    finally:
        for g in groups:
            g.close()
        torch.cuda.set_device(0)


def _mg_worker(rank, world, port, mode, q):
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
        M, N = 4096, 8192
        A = ob.gen_matrix(M, N, 0.98, 420)
        x = ob.gen_vector(M, 0.5, 421)
        y64, s = ob.sgemv_dense_f64(A, x)
        a, b = rank * N // world, (rank + 1) * N // world
        if mode == "ipc":                                 # library-owned blocks, CUDA IPC handles over the caller's transport
            g = S.Group(M, N, rank, world)
            handles = [None] * world
            dist.all_gather_object(handles, g.ipc_handle())
            g.connect_ipc(handles)
        else:                                             # symmetric memory (+ multicast alias when NVLS is there)
            import torch.distributed._symmetric_memory as symm
            block = symm.empty(S.Group.block_bytes(N) // 4, dtype=torch.float32, device=torch.device("cuda", rank))
            block.zero_()
            hdl = symm.rendezvous(block, dist.group.WORLD)
            g = S.Group(M, N, rank, world, block.data_ptr())
            g.connect_ptrs([int(p) for p in hdl.buffer_ptrs], int(getattr(hdl, "multicast_ptr", 0) or 0))
        torch.cuda.synchronize()
        dist.barrier()
        half = (b - a) // 2
        g.add(S.Plan.from_dense("awsp", A[:, a:a + half], chunk_mode=4), a)
        g.add(S.Plan.from_dense("tcsr", A[:, a + half:b]), a + half)
        ys = [g.run_host(x) for _ in range(4)]            # the FULL y on every rank, four calls (both buffers twice)
        g.status()
        err = float(np.max(np.abs(ys[-1].astype(np.float64) - y64) / (s + 1e-30)))
        q.put((rank, err, all(v.tobytes() == ys[0].tobytes() for v in ys), ys[-1].tobytes()))
        dist.barrier()
        g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ipc", "symm"])
def test_group_two_ranks(mode):
    """One process per GPU: every rank ends up with the full y, bit-identical on both ranks and within
    the parity gate of the oracle — the fused epilogue plus the in-kernel arrival, no NCCL on the data path."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mg_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(e <= 1e-5 and stable for _, e, stable, _ in res), [(r, e, st) for r, e, st, _ in res]
    assert res[0][3] == res[1][3], "the two ranks hold different y"

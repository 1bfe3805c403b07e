#!/usr/bin/env python
"""Development tool: per-call time of back-to-back launches in a stream vs the same calls
replayed from one CUDA graph (is the ~2 us quantisation of stream launches real work?)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spmv_test_b200 as S
from spmv_test_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
M, N, sa, sx = synth.CONFIGS[cfg]
A = synth.gen_matrix(M, N, sa)
x = synth.gen_vector(M, sx)
st = torch.cuda.Stream()
for v in ("wsp", "asp", "awsp", "tcsr"):
    plan = S.Plan.from_dense(v, A)
    plans = bench.make_copies(plan)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(N, device="cuda")
    ms = bench.time_loop(torch, plans, dx, dy, 400, 20, st, graph=False)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        for i in range(8):
            plans[i % len(plans)].run(dx, dy, st.cuda_stream)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(100):
                plans[i % len(plans)].run(dx, dy, torch.cuda.current_stream().cuda_stream)
        g.replay(); g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for _ in range(4):
            g.replay()
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
    print(f"{v:5s} {cfg}: stream {ms * 1e3 / 400:7.3f} us/call   graph {e0.elapsed_time(e1) * 1e3 / 400:7.3f} us/call", flush=True)
    for p in plans:
        p.close()

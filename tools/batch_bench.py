#!/usr/bin/env python
"""Development tool: batched wsp (spmv_run_batch) on config 2 — µs per call and effective GB/s
per batch (algorithmic bytes of `batch` single-vector calls / time)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import numpy as np
import spmv_test_b200 as S
from spmv_test_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
M, N, sa, sx = synth.CONFIGS[cfg]
A = synth.gen_matrix(M, N, sa)
st = torch.cuda.Stream()
variant = sys.argv[2] if len(sys.argv) > 2 else "wsp"
plan = S.Plan.from_dense(variant, A)
plans = bench.make_copies(plan)
for batch in (1, 2, 4, 8):
    X = np.stack([synth.gen_vector(M, sx, seed=10 + b) for b in range(batch)])
    dX = torch.from_numpy(X).cuda()
    dY = torch.zeros((batch, N), device="cuda")
    alg = sum(plan.traffic(X[b])[0] for b in range(batch))
    n = len(plans)
    ms, _ = bench.timed_steps(torch, lambda i, cs: plans[i % n].run_batch(dX, dY, cs), 400, 20, st)
    us = ms * 1e3 / 400
    print(f"{variant} {cfg} batch {batch}: {us:8.3f} us per batched call, {us / batch:7.3f} us per vector, "
          f"{alg / (us * 1e-6) / 1e9:8.1f} GB/s effective", flush=True)

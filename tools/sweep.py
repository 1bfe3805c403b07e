#!/usr/bin/env python
"""Development tool: time one variant/config over a grid of plan options (needs a GPU).
    python tools/sweep.py awsp c2 warps_per_col=4,8 row_splits=8,16,32"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spmv_test_b200 as S
from spmv_test_b200 import synth

variant, cfg = sys.argv[1], sys.argv[2]
grid = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    grid[k] = [int(t) for t in v.split(",")]
M, N, sa, sx = synth.CONFIGS[cfg]
A = synth.gen_matrix(M, N, sa)
x = synth.gen_vector(M, sx)
stream = torch.cuda.Stream()
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    opts = dict(zip(keys, combo))
    try:
        r, plans, _, _ = bench.measure_variant(torch, S, variant, lambda v: S.Plan.from_dense(v, A, **opts), x, 400, 20, stream)
        print(variant, cfg, opts, "->", r["us_per_call"], "us  phys", r["phys_GBps"], "GB/s  grid", r["grid"], flush=True)
        for p in plans:
            p.close()
    except Exception as e:
        print(variant, cfg, opts, "failed:", e, flush=True)

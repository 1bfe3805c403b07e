#!/usr/bin/env python
"""Development tool: one GPU's config-5 slab (65536 x 131072, 1 %) over a grid of plan options.
    python tools/c5_slab.py chunk_mode=1,2 slab_cols=2048,4096"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spmv_test_b200 as S
from spmv_test_b200 import synth

variant = "awsp"
grid = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "variant":
        variant = v
    else:
        grid[k] = [int(t) for t in v.split(",")]
st = torch.cuda.Stream()
cp, ri, va = synth.bernoulli_csc(bench.C5_M, bench.C5_SLAB_N, bench.C5_DENSITY, 5000)
x = synth.gen_vector(bench.C5_M, bench.C5_SX)
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    opts = dict(zip(keys, combo))
    try:
        r, pl, _, _ = bench.measure_variant(torch, S, variant, lambda v: S.Plan.from_csc(v, bench.C5_M, bench.C5_SLAB_N, cp, ri, va, **opts),
                                            x, 100, 5, st)
        print("c5slab", variant, opts, r["us_per_call"], "us  phys", r["phys_GBps"], " alg", r["eff_GBps"], "GB/s  grid", r["grid"],
              "W", r["slab_cols"], flush=True)
        for p in pl:
            p.close()
    except Exception as e:
        print("c5slab", variant, opts, "failed:", e, flush=True)

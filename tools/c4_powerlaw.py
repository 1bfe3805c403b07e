#!/usr/bin/env python
"""Development tool: config 4 (1M x 1M power-law columns, wsp, dense x) timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import spmv_test_b200 as S
from spmv_test_b200 import synth

st = torch.cuda.Stream()
cp, ri, va = synth.powerlaw_csc(1 << 20, 1 << 20, seed=42)
x = synth.gen_vector(1 << 20, 0.0, seed=7)
r, pl, _, _ = bench.measure_variant(torch, S, "wsp", lambda v: S.Plan.from_csc(v, 1 << 20, 1 << 20, cp, ri, va), x, 100, 5, st)
print("c4", os.environ.get("SPMV_B200_LIB", "default"), r["us_per_call"], "us phys", r["phys_GBps"], "alg", r["eff_GBps"], r["grid"], r["kernels_per_call"])

#!/usr/bin/env python
"""Small all-variant run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_test_b200 as S
from spmv_test_b200 import synth

for (M, N, sa, sx, opts) in [(256, 512, 0.7, 0.5, {}), (160, 256, 0.97, 0.5, dict(chunk_mode=2)), (96, 1024, 0.5, 0.0, dict(row_splits=3))]:
    A = synth.gen_matrix(M, N, sa, seed=M)
    x = synth.gen_vector(M, sx, seed=N)
    ref = x.astype(np.float64) @ A.astype(np.float64)
    for v in ("wsp", "asp", "awsp", "tcsr"):
        o = opts if v in ("awsp", "tcsr") or "row_splits" in opts else {}
        if v == "wsp":
            o = {}
        with S.Plan.from_dense(v, A, **o) as p:
            y = p.run_host(x)
            assert np.allclose(y, ref, atol=1e-3), (v, M, N)
print("sanitize_small: ok")

#!/usr/bin/env python
"""Development tool: per-warp timeline of the panel kernel (needs tools/trace_build.sh and a GPU).
    SPMV_B200_LIB=spmv_test_b200/lib/libspmv_b200_trace.so python tools/trace_panel.py [variant] [config]
Stamps: 0 start, 1 first metadata ready, 2 ring filled (issued), 3 first chunk landed,
4 stream done, 5 after CTA barrier, 6 partial written, 7 end (after split reduce)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_test_b200 as S
from spmv_test_b200 import synth

variant = sys.argv[1] if len(sys.argv) > 1 else "awsp"
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
opts = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    opts[k] = int(v)
M, N, sa, sx = synth.CONFIGS[cfg]
A = synth.gen_matrix(M, N, sa)
x = synth.gen_vector(M, sx)
plan = S.Plan.from_dense(variant, A, **opts)
plans = [plan] + [plan.clone() for _ in range(3)]
dx = torch.from_numpy(x).cuda()
dy = torch.zeros(N, device="cuda")
st = torch.cuda.Stream()
info = plan.info()
print(info)
nw = info["grid_x"] * info["grid_y"] * 8
for i in range(9):
    plans[i % 4].run(dx, dy, st.cuda_stream)
st.synchronize()
buf = np.zeros(nw * 10, np.uint64)
L = S.lib()
L.spmv_trace_read.argtypes = [C.c_void_p, C.c_int64]
assert L.spmv_trace_read(C.c_void_p(buf.ctypes.data), buf.size) == 0
raw = buf.reshape(nw, 10).astype(np.int64)
smid = raw[:, 8]
t = raw[:, :8]
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["start", "meta", "issued", "first", "streamed", "cta_bar", "partial", "end"]
for k, n in enumerate(names):
    c = t[:, k]
    print(f"{n:9s} min {c.min():7.2f}  p10 {np.percentile(c, 10):7.2f}  med {np.median(c):7.2f}  p90 {np.percentile(c, 90):7.2f}  max {c.max():7.2f} us")
d = np.diff(t, axis=1)
print("phase medians (us):", {names[k + 1]: round(float(np.median(d[:, k])), 2) for k in range(7)})
print("phase p90 (us):    ", {names[k + 1]: round(float(np.percentile(d[:, k], 90)), 2) for k in range(7)})

# stragglers: the warps that finish streaming last
order = np.argsort(-t[:, 4])[:12]
for w in order:
    print(f"warp {w:5d} cta {w // 8:4d} sm {smid[w]:3d}  " + " ".join(f"{names[k]}={t[w, k]:.1f}" for k in range(8)))
per_sm = {}
for w in range(nw):
    per_sm.setdefault(int(smid[w]), []).append(t[w, 4])
ends = sorted((max(v), k, len(v)) for k, v in per_sm.items())
print("SMs by last 'streamed':", [(k, n, round(e, 1)) for e, k, n in ends[-8:]], " fastest:", [(k, n, round(e, 1)) for e, k, n in ends[:4]])

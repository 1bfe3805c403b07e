#!/usr/bin/env python
"""Development tool: per-warp timeline of panel_rs_kernel + panel_rs_reduce_kernel (needs tools/trace_build.sh and a GPU).
    SPMV_B200_LIB=spmv_test_b200/lib/libspmv_b200_trace.so python tools/trace_rs.py [variant] [config]
Main kernel stamps: 0 start, 1 after the grid dependency wait, 2 first list built, 3 first slots consumed,
4 streamed (loop left), 5 drained, 6 after the CTA barrier, 7 partial row written.  Reduce kernel (warps from 32768 on):
0 start, 1 after the wait, 2 end.  Times are relative to the main kernel's first start of the LAST of 9 back-to-back calls."""
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
M, N, sa, sx = synth.CONFIGS[cfg]
A = synth.gen_matrix(M, N, sa)
x = synth.gen_vector(M, sx)
plan = S.Plan.from_dense(variant, A)
plans = [plan] + [plan.clone() for _ in range(7)]       # 8 x 53 MB: every call finds its matrix cold in L2
dx = torch.from_numpy(x).cuda()
dy = torch.zeros(N, device="cuda")
st = torch.cuda.Stream()
info = plan.info()
print(info)
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(st):
    for i in range(8):
        plans[i % 8].run(dx, dy, st.cuda_stream)
    st.synchronize()
    with torch.cuda.graph(g, stream=st):
        for i in range(9):
            plans[i % 8].run(dx, dy, st.cuda_stream)
    g.replay(); g.replay()
st.synchronize()
torch.cuda.synchronize()
buf = np.zeros(65536 * 10, np.uint64)
L = S.lib()
L.spmv_trace_read.argtypes = [C.c_void_p, C.c_int64]
assert L.spmv_trace_read(C.c_void_p(buf.ctypes.data), buf.size) == 0
raw = buf.reshape(65536, 10).astype(np.int64)
nw = info["grid_x"] * info["grid_y"] * 8
main = raw[:nw]
red = raw[32768:]
red = red[red[:, 0] > 0][:, :3]
t0 = main[:, 0].min()
t = (main[:, :8] - t0) / 1e3
names = ["start", "waited", "list", "first", "streamed", "drained", "cta_bar", "partial"]
for k, n in enumerate(names):
    c = t[:, k]
    print(f"{n:9s} min {c.min():7.2f}  p10 {np.percentile(c, 10):7.2f}  med {np.median(c):7.2f}  p90 {np.percentile(c, 90):7.2f}  max {c.max():7.2f} us")
r = (red - t0) / 1e3
for k, n in enumerate(["r_start", "r_waited", "r_end"]):
    c = r[:, k]
    print(f"{n:9s} min {c.min():7.2f}  med {np.median(c):7.2f}  max {c.max():7.2f} us   ({len(c)} warps)")
smid = main[:, 8]
per_sm = {}
for w in range(nw):
    per_sm.setdefault(int(smid[w]), []).append(t[w, 4])
ends = sorted((max(v), k, len(v)) for k, v in per_sm.items())
print("SMs by last 'streamed':", [(k, n, round(e, 1)) for e, k, n in ends[-8:]], " fastest:", [(k, n, round(e, 1)) for e, k, n in ends[:4]])
print("warps per SM:", sorted(set(len(v) for v in per_sm.values())))

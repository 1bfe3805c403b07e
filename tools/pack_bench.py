"""Plan-creation wall time: host packers vs device packers.
usage: python tools/pack_bench.py [config ...]   (default: c2)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import spmv_test_b200 as S
from spmv_test_b200 import synth

cfgs = sys.argv[1:] or ["c2"]
for name in cfgs:
    M, N, sA, _ = synth.CONFIGS[name]
    A = synth.gen_matrix(M, N, sA, 1234)
    dA = torch.from_numpy(A).cuda()
    torch.cuda.synchronize()
    for v in ("wsp", "awsp", "tcsr", "asp"):
        res = {}
        for label, make in (("host", lambda: S.Plan.from_dense(v, A, pack_mode="host")),
                            ("device(host A)", lambda: S.Plan.from_dense(v, A, pack_mode="device")),
                            ("device(device A)", lambda: S.Plan.from_dense_device(v, dA))):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                p = make()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
                p.close()
            res[label] = best * 1e3
        print(name, v, {k: round(t, 2) for k, t in res.items()}, "ms", flush=True)

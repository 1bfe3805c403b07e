"""bench.py's CPU-side pieces: the algorithmic-byte accounting and the
reference arm's JSON line (the driver runs `bench.py --impl reference` next to the GPU arm)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_alg_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(3)
    M, N = 96, 64
    A = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    A[rng.random((M, N)) < 0.7] = 0
    x = rng.uniform(-1, 1, M).astype(np.float32)
    x[::2] = 0
    nnz = np.count_nonzero(A)
    nnz_t = np.count_nonzero(A[x != 0])
    vec = 4 * M + 4 * N
    assert bench.step_alg_bytes(A, x, ["wsp"]) == 8 * nnz + 4 * (N + 1) + vec
    assert bench.step_alg_bytes(A, x, ["awsp"]) == 8 * nnz_t + 4 * (N + 1) + vec
    assert bench.step_alg_bytes(A, x, ["tcsr"]) == bench.step_alg_bytes(A, x, ["awsp"])
    assert bench.step_alg_bytes(A, x, ["asp"]) == 4 * np.count_nonzero(x) * N + vec
    assert bench.step_alg_bytes(A, x, ["wsp", "asp", "awsp"]) == sum(bench.step_alg_bytes(A, x, [v]) for v in ("wsp", "asp", "awsp"))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["unit"] == "GB/s" and line["higher_is_better"] is True and line["value"] > 0 and line["scaling"] == "strong"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == line["value"]
    assert cb["dense_reference"]["kind"] in ("reference", "port") and cb["dense_reference"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config 5" in line["config"]["workload"] and line["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.c5_config()            # the same config dict as the GPU arm (modulo its l2 / launch notes)


def test_config5_sampling_and_alg_bytes():
    sys.path.insert(0, ROOT)
    import bench
    from spmv_test_b200 import synth
    cp, ri, va = synth.bernoulli_csc(512, bench.C5_SLAB_N, 0.01, seed=5000)
    cols, scp, sri, sva = bench.c5_sample(cp, ri, va, 0, n_cols=64)
    assert cols.size == 64 and np.all(np.diff(cols) > 0) and scp[-1] == sri.size == sva.size
    for k in (0, 17, 63):
        c = cols[k]
        assert np.array_equal(sri[scp[k]:scp[k + 1]], ri[cp[c]:cp[c + 1]]) and np.array_equal(sva[scp[k]:scp[k + 1]], va[cp[c]:cp[c + 1]])
    N = bench.C5_SLAB_N * bench.C5_SLABS
    assert bench.c5_alg_bytes(1000) == 8 * 1000 + 4 * (N + 1) + 4 * bench.C5_M + 4 * N

#!/usr/bin/env python
"""bench.py — sparse SGEMV Y = x·A on B200: µs/call and effective HBM GB/s vs the roofline.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                    (the CPU path, rank 0)

Workload at EVERY N (strong scaling): BASELINE config 5 — the fixed matrix A 65536 x 1048576, 99 %
sparse, built directly in sparse form as eight 131072-column slabs (seeds 5000..5007), x 50 %
activation-sparse, awsp variant in the row-strip form (chunk_mode 4).  Rank r owns the slabs
[8r/N, 8(r+1)/N); a step is y = x·A for the whole matrix: every rank runs its slabs back to back,
their epilogues store each slice of y into every rank's copy of y (fused all-gather) and a
one-warp arrival kernel joins the ranks (spmv_mg_run).  At N = 1 the line also carries every
other BASELINE config per variant ("configs"), each checked against the oracle.

value   = algorithmic bytes (8*nnz_touched + 4(N+1) + 4M + 4N) of the whole matrix divided
          by the device time of a step (CUDA events on the launching stream, max over ranks), GB/s.
L2      config 5: a rank's slabs are 0.7 GB (N = 8) .. 5.6 GB (N = 1), far beyond the 126 MB L2.
          Small configs: every timed loop rotates over clones of the packed matrix whose total
          size exceeds 2.5x L2; a hot-L2 figure (one copy) is reported separately.
parity  every timed plan is first checked against the CPU oracle (tests/parity.py gate); a
          failure makes the exit code non-zero.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sparse SGEMV Y=xA effective HBM throughput (algorithmic bytes / device time)"
UNIT = "GB/s"
L2_BYTES = 126e6
HEADLINE = "awsp"
C5_M, C5_SLAB_N, C5_SLABS, C5_DENSITY, C5_SX = 65536, 131072, 8, 0.01, 0.5
C5_CHUNK_MODE = 4                       # row strips (strips.cu); 3 = lane-owned blocks, 0 = multi-row panel form
C5_CHECK_COLS = 1024                    # oracle columns sampled per slab
PARITY_FAILED = []


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram bytes per launch from the committed ncu capture, if there is one for this kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def oracle_bindings():
    """The checker (test infrastructure): only the parity checks and the CPU baseline legs use it."""
    t = os.path.join(ROOT, "tests")
    if t not in sys.path:
        sys.path.insert(0, t)
    import oracle_bindings as ob
    import parity
    return ob, parity


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML while the timed regions run."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz, self.ok = index, [], 0, False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        busy = [s for s in self.samples if s > 0]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "samples": len(busy)}


# ------------------------------------------------------------------------------------------------
def make_copies(plan, want_bytes=2.5 * L2_BYTES, max_copies=96):
    """Clones of the resident matrix so that a rotation over them defeats the L2."""
    each = max(1, plan.info()["device_bytes"])
    n = int(min(max_copies, max(1, -(-want_bytes // each))))
    return [plan] + [plan.clone() for _ in range(n - 1)]


GRAPH_STEPS = 100   # steps captured per CUDA graph at most
GRAPH = True        # --no-graph clears it
CHUNKS = 4          # the timed steps run as this many separately timed chunks (min / median)


def timed_steps(torch, step, steps, warmup, stream, graph=True, barrier=None):
    """Device time of exactly `steps` calls of step(i, cuda_stream) on `stream`: CUDA events on that
    stream, device-wide synchronize (and `barrier()`, across ranks) on both sides.  The steps run as
    up to CHUNKS chunks, each between its own pair of events, so besides the total (first event to
    last) there is a per-step time for every chunk: min / median.  With graph=True a chunk is a
    CUDA graph replay (kernels launched one by one into a stream start on a ~2 us dispatch cadence
    on this driver; a graph runs them back to back).  Chunks are an even number of steps (the
    multi-GPU y buffers alternate).  Returns (total_ms, [us per step of every chunk])."""
    n_chunks = max(1, min(CHUNKS, steps // 2))
    per = steps // n_chunks
    if per > 1 and per % 2:
        per -= 1
    per = max(1, min(per, GRAPH_STEPS if graph else per))
    sizes = []
    left = steps
    while left > 0:
        n = min(per, left)
        sizes.append(n)
        left -= n
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(sizes) + 1)]
    with torch.cuda.stream(stream):
        for i in range(warmup):
            step(i, stream.cuda_stream)
        stream.synchronize()
        graphs = {}
        if graph:
            for n in sorted(set(sizes)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    cs = torch.cuda.current_stream().cuda_stream
                    for i in range(n):
                        step(i, cs)
                graphs[n] = g
                g.replay()                                 # untimed: the first replay uploads the graph
            stream.synchronize()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        ev[0].record(stream)
        done = 0
        for k, n in enumerate(sizes):
            if graph:
                graphs[n].replay()
            else:
                for i in range(n):
                    step(done + i, stream.cuda_stream)
            done += n
            ev[k + 1].record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if barrier:
            barrier()
    total = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) * 1e3 / n for k, n in enumerate(sizes)]
    return total, per_step


def stats(per_step):
    return {"min": round(min(per_step), 3), "median": round(statistics.median(per_step), 3), "chunks": len(per_step)}


def record_parity(what, fn):
    """Runs a parity check; a failure is recorded (exit code != 0 at the end) instead of losing the line."""
    try:
        fn()
        return "ok"
    except AssertionError as e:
        PARITY_FAILED.append(f"{what}: {e}")
        return "FAILED: " + str(e)[:160]


def measure_variant(torch, S, variant, build, x, steps, warmup, stream, check=None, hot=False, e2e=False):
    """Pack, check against the oracle, clone, time.  `build(variant)` returns a Plan;
    `check(variant, y)` raises AssertionError on a parity failure."""
    t0 = time.perf_counter()
    plan = build(variant)
    pack_s = time.perf_counter() - t0
    info = plan.info()
    alg, phys, nnz_t = plan.traffic(x)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(info["N"], dtype=torch.float32, device="cuda")
    parity = None
    if check is not None:
        plan.run(dx, dy, stream.cuda_stream)
        stream.synchronize()
        y = dy.cpu().numpy()
        parity = record_parity(f"{variant} {info['M']}x{info['N']}", lambda: check(variant, y))
    plans = make_copies(plan)
    n = len(plans)
    ms, per = timed_steps(torch, lambda i, cs: plans[i % n].run(dx, dy, cs), steps, warmup, stream, graph=GRAPH)
    us = ms * 1e3 / steps
    res = {"us_per_call": round(us, 3), "us_min": stats(per)["min"], "us_median": stats(per)["median"],
           "alg_MB": round(alg / 1e6, 3), "phys_MB": round(phys / 1e6, 3),
           "eff_GBps": round(alg / (us * 1e-6) / 1e9, 1), "phys_GBps": round(phys / (us * 1e-6) / 1e9, 1),
           "nnz_touched": nnz_t, "l2_copies": n, "resident_MB": round(info["device_bytes"] / 1e6, 1),
           "grid": [info["grid_x"], info["grid_y"]], "kernels_per_call": info["kernels_per_run"],
           "slab_cols": info["slab_cols"], "row_splits": info["row_splits"], "pack_s": round(pack_s, 2)}
    if parity is not None:
        res["parity"] = parity
    if e2e:                                               # the reference launcher's per-call part: pinned host x in, y out, synchronise
        hx = torch.from_numpy(x).pin_memory()
        hy = torch.empty(info["N"], dtype=torch.float32).pin_memory()
        n_e = max(20, min(steps, 400))
        for _ in range(5):
            plan.run_host_ptr(hx.data_ptr(), hy.data_ptr())
        t0 = time.perf_counter()
        for _ in range(n_e):
            plan.run_host_ptr(hx.data_ptr(), hy.data_ptr())
        res["e2e_us_per_call"] = round((time.perf_counter() - t0) / n_e * 1e6, 2)
        res["e2e_bytes"] = {"h2d": int(hx.numel() * 4), "d2h": int(hy.numel() * 4)}
    if hot:                                               # the same call on ONE resident copy: L2-warm when it fits
        ms_h, per_h = timed_steps(torch, lambda i, cs: plan.run(dx, dy, cs), steps, warmup, stream, graph=GRAPH)
        res["hot_l2_us_per_call"] = round(ms_h * 1e3 / steps, 3)
    return res, plans, (dx, dy), (alg, phys)


# ------------------------------------------------------------------------------------------------
def step_alg_bytes(A, x, names):
    """Algorithmic bytes of one call of each variant in `names`."""
    M, N = A.shape
    nnz = int(np.count_nonzero(A))
    act = x != 0
    nnz_t = int(np.count_nonzero(A[act]))
    mnz = int(np.count_nonzero(act))
    vec = 4.0 * M + 4.0 * N
    per = {"wsp": 8.0 * nnz + 4 * (N + 1) + vec, "tcsr": 8.0 * nnz_t + 4 * (N + 1) + vec,
           "awsp": 8.0 * nnz_t + 4 * (N + 1) + vec, "asp": 4.0 * mnz * N + vec}
    return sum(per[v] for v in names)


def c5_alg_bytes(nnz_touched_total):
    N = C5_SLAB_N * C5_SLABS
    return 8.0 * nnz_touched_total + 4.0 * (N + 1) + 4.0 * C5_M + 4.0 * N


def c5_config():
    """The same dict in the GPU arm and the reference arm (the driver compares them)."""
    return {"workload": "BASELINE config 5 (strong scaling): A 65536x1048576 fp32, 99% sparse, built in sparse form as eight "
                        "131072-column slabs (seeds 5000..5007), x 50% activation-sparse (seed 4321); awsp, row strips "
                        "(chunk_mode 4); one step = y = x.A for the whole matrix, rank r owns slabs [8r/N, 8(r+1)/N), "
                        "fused all-gather of Y + in-kernel arrival",
            "M": C5_M, "N": C5_SLAB_N * C5_SLABS, "slabs": C5_SLABS, "weight_sparsity": 1.0 - C5_DENSITY,
            "activation_sparsity": C5_SX, "variant": HEADLINE, "chunk_mode": C5_CHUNK_MODE,
            "l2": "inputs larger than L2: a rank's resident slabs are 0.7 GB (N = 8) to 5.6 GB (N = 1) against 126 MB of L2, and "
                  "every step streams all of them once; no rotation or flush needed",
            "launch": "GPU arm: CUDA graphs of the steps, timed in chunks (see 'timing'); reference arm: host loop"}


def c5_slab_csc(synth, g):
    return synth.bernoulli_csc(C5_M, C5_SLAB_N, C5_DENSITY, seed=5000 + g)


def c5_sample(col_ptr, row_idx, vals, g, n_cols=C5_CHECK_COLS):
    """A seeded sample of a slab's columns as a small CSR(A^T) of its own (for the oracle)."""
    rng = np.random.default_rng(900 + g)
    cols = np.sort(rng.choice(C5_SLAB_N, size=n_cols, replace=False))
    lens = (col_ptr[cols + 1] - col_ptr[cols]).astype(np.int64)
    cp = np.zeros(n_cols + 1, np.int64)
    np.cumsum(lens, out=cp[1:])
    take = np.concatenate([np.arange(col_ptr[c], col_ptr[c + 1]) for c in cols]) if n_cols else np.zeros(0, np.int64)
    return cols, cp, row_idx[take].copy(), vals[take].copy()


def check_sampled(y_slab, sample, x, what):
    """y of one slab against the oracle on its sampled columns (fp32 sequential + fp64, parity.py gate)."""
    ob, parity = oracle_bindings()
    cols, cp, ri, va = sample
    y32 = ob.csc_gemv(cols.size, cp, ri, va, x)
    xa = x.astype(np.float64)[ri] * va.astype(np.float64)
    seg = np.repeat(np.arange(cols.size), np.diff(cp))
    y64 = np.bincount(seg, weights=xa, minlength=cols.size)
    s = np.bincount(seg, weights=np.abs(xa), minlength=cols.size)
    parity.check_y(y_slab[cols], y32, y64, s, what)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
def cpu_baselines(A, x, col_csc, alg_bytes, budget_s=6.0):
    """BASELINE.md §5 beside the GPU number, same inputs: (i) the reference's dense SgemvCPU (tester.cpp:36-45,
    from oracle/_ref when built, else the oracle's restatement), 1 thread, dense A only; (ii) the CSR(A^T)
    restatement of csr_naive.cu:14-22 on 1 thread and on all host cores (OpenMP).  1 warm-up + best of
    up to 5, bounded by `budget_s` per leg.  Returns a dict of GB/s of algorithmic bytes and ms per call."""
    ob, _ = oracle_bindings()
    cores = os.cpu_count() or 1
    out = {"nproc": cores}

    def best_of(fn):
        fn()
        best, t_all = None, time.perf_counter()
        for _ in range(5):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            if time.perf_counter() - t_all > budget_s:
                break
        return best
    if A is not None:
        M = A.shape[0]
        rows = M if M * A.shape[1] <= 4096 * 4096 else max(32, (M // 8) // 32 * 32)   # bounded sample of the rows
        Ar, xr = np.ascontiguousarray(A[:rows]), np.ascontiguousarray(x[:rows])
        fn, kind = (ob.ref_sgemv_cpu, "reference") if ob.have_ref_cpu() else (ob.sgemv_dense, "port")
        dt = best_of(lambda: fn(Ar, xr)) * (M / rows)
        out["dense_1t"] = {"ms_per_call": round(dt * 1e3, 3), "GBps": round(alg_bytes / dt / 1e9, 4), "kind": kind,
                           "sample": f"first {rows} of {M} rows, scaled"}
    N, cp, ri, va = col_csc
    for name, th in (("csr_1t", 1), ("csr_all_cores", cores)):
        dt = best_of(lambda: ob.csc_gemv(N, cp, ri, va, x, threads=th))
        out[name] = {"ms_per_call": round(dt * 1e3, 3), "GBps": round(alg_bytes / dt / 1e9, 4), "threads": th}
    return out


def run_reference_arm(args):
    """The CPU arm on the headline workload (config 5).  The reference's own CPU path is the dense
    single-thread SgemvCPU (tester.cpp:36-45), which cannot hold this matrix (256 GB dense); the
    arm therefore times the reference's sparse algorithm — csr_naive.cu:14-22 restated for the CPU
    in the oracle (CSR of A^T, sequential per output) — with OpenMP on all host cores, on a bounded
    sample: ONE of the eight slabs per step (outputs are independent: time is linear in slabs).
    The dense SgemvCPU on a 65536 x 1024 block of the same slab is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from spmv_test_b200 import synth
    ob, _ = oracle_bindings()
    cores = os.cpu_count() or 1
    cp, ri, va = c5_slab_csc(synth, 0)
    x = synth.gen_vector(C5_M, C5_SX, seed=4321)
    nnz_t = int(np.count_nonzero(x[ri] != 0))
    alg_slab = c5_alg_bytes(nnz_t * C5_SLABS) / C5_SLABS
    for _ in range(args.warmup):
        ob.csc_gemv(C5_SLAB_N, cp, ri, va, x, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.csc_gemv(C5_SLAB_N, cp, ri, va, x, threads=cores)
    dt = (time.perf_counter() - t0) / max(1, args.steps)        # per slab
    gbps = alg_slab / dt / 1e9
    # the reference's dense loop on a dense block of the slab's first 1024 columns
    cols = 1024
    A = np.zeros((C5_M, cols), np.float32)
    A[ri[: cp[cols]], np.repeat(np.arange(cols), np.diff(cp[: cols + 1]))] = va[: cp[cols]]
    fn, kind = (ob.ref_sgemv_cpu, "reference") if ob.have_ref_cpu() else (ob.sgemv_dense, "port")
    t0 = time.perf_counter()
    fn(A, x)
    dt_dense = time.perf_counter() - t0
    alg_block = 8.0 * float(np.count_nonzero(A[x != 0.0])) + 4.0 * (cols + 1) + 4.0 * C5_M / (C5_SLAB_N / cols) + 4.0 * cols
    sample = (f"{args.steps} steps, each ONE of the 8 slabs (65536x131072, {ri.size} non-zeros) through the oracle's CSR(A^T) "
              f"restatement of csr_naive.cu:14-22 with OpenMP on {cores} host threads: {dt * 1e3:.1f} ms per slab "
              f"(whole matrix ~{dt * C5_SLABS * 1e3:.0f} ms)")
    line = {"impl": "reference", "metric": METRIC, "value": round(gbps, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * C5_SLABS * 1e3, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": c5_config(),
            "cpu_baseline": {"value": round(gbps, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "dense_reference": {"kind": kind, "cores": 1, "GBps": round(alg_block / dt_dense / 1e9, 4),
                                                 "sample": f"SgemvCPU (tester.cpp:36-45) on the dense 65536x{cols} block of the "
                                                           f"slab's first columns: {dt_dense * 1e3:.0f} ms (the matrix is "
                                                           f"{C5_SLAB_N * C5_SLABS // cols} such blocks)"}},
            "e2e": {"value": round(gbps, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def build_c5_rank(S, synth, slab_ids, chunk_mode):
    """This rank's slabs: generate (a small pool of host threads), sample for the oracle, pack."""
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(4, max(1, len(slab_ids)))) as ex:
        cscs = list(ex.map(lambda g: c5_slab_csc(synth, g), slab_ids))
    gen_s = time.perf_counter() - t0
    plans, samples, nnz = [], [], []
    t0 = time.perf_counter()
    import torch
    for g, (cp, ri, va) in zip(slab_ids, cscs):
        samples.append(c5_sample(cp, ri, va, g))
        if chunk_mode == 4:                               # row strips: packed by kernels from the CSR(A^T) arrays in HBM
            d = [torch.from_numpy(a).cuda() for a in (cp, ri, va)]
            plans.append(S.Plan.from_csc_device(HEADLINE, C5_M, C5_SLAB_N, *d, chunk_mode=chunk_mode))
            del d
        else:
            plans.append(S.Plan.from_csc(HEADLINE, C5_M, C5_SLAB_N, cp, ri, va, chunk_mode=chunk_mode))
        nnz.append(int(ri.size))
    cscs.clear()
    return plans, samples, nnz, gen_s, time.perf_counter() - t0


def gpu_comparators(torch, A, x, y_ref, stream):
    """The reference's own kernels recompiled for sm_100a (oracle/_ref/libspmv_ref_gpu.so: awsp_ref.cu:6-185,
    wsp.cu:59-138, asp.cu:116-211, csr_naive.cu, cublas.cu:33) on the 4096x4096 harness shape, as the
    reference's launchers run them (one cold launch each, their own TIME_KERNEL bracket), next to
    cublasSgemv through torch.  Test infrastructure: built only where /root/reference exists."""
    ob, _ = oracle_bindings()
    out = {}
    if not ob.have_ref_gpu():
        return {"unavailable": "oracle/_ref/libspmv_ref_gpu.so was not built (no /root/reference at build time)"}
    for name, ver in (("awsp_ref", 0), ("awsp", 2), ("wsp", 1), ("asp", 2), ("csr_naive", 0), ("csr_tiling", 0), ("cublas", 0)):
        try:
            best = None
            for _ in range(3):
                y, ms = ob.ref_gpu_gemv(name, A, x, version=ver)
                best = ms if best is None else min(best, ms)
            out[f"{name}_v{ver}"] = {"us_per_call": round(best * 1e3, 2), "max_abs_diff_vs_oracle": float(np.max(np.abs(y - y_ref)))}
        except Exception as e:
            out[name] = {"error": str(e)[:120]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline only (used under ncu)")
    ap.add_argument("--no-aux", action="store_true", help="skip the CPU-baseline / comparator / batched legs")
    ap.add_argument("--no-graph", action="store_true", help="launch every call into the stream instead of replaying CUDA graphs")
    ap.add_argument("--chunk-mode", type=int, default=C5_CHUNK_MODE, help="config-5 form: 4 row strips, 3 lane-owned blocks, 0 multi-row")
    ap.add_argument("--join", default="arrive", choices=["arrive", "nccl"], help="N > 1: in-kernel arrival (default) or NCCL all_gather")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 10 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_reference_arm(args)
        return
    global GRAPH
    GRAPH = not args.no_graph
    args.steps = 200 if args.steps is None else args.steps
    args.warmup = 20 if args.warmup is None else max(3, args.warmup)

    import torch
    import spmv_test_b200 as S
    from spmv_test_b200 import synth
    if not torch.cuda.is_available() or S.lib().spmv_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > C5_SLABS:
        raise SystemExit(f"bench.py: the config-5 matrix has {C5_SLABS} column slabs; run with at most {C5_SLABS} GPUs")
    stream = torch.cuda.Stream()
    peak, peak_src = peaks()
    sampler = ClockSampler(local)
    sampler.start()
    extra = {}
    barrier = (lambda: dist.barrier()) if dist is not None else None

    # ---------------- the headline: config 5, this rank's slabs ------------------------------------
    slab_ids = list(range(rank * C5_SLABS // world, (rank + 1) * C5_SLABS // world))
    plans, samples, nnz, gen_s, pack_s = build_c5_rank(S, synth, slab_ids, args.chunk_mode)
    x = synth.gen_vector(C5_M, C5_SX, seed=4321)
    dx = torch.from_numpy(x).cuda()
    N_total = C5_SLAB_N * C5_SLABS
    tr = [p.traffic(x) for p in plans]
    alg_rank, phys_rank, touched_rank = sum(t[0] for t in tr), sum(t[1] for t in tr), sum(t[2] for t in tr)

    # the shared block: two copies of the full y + arrival flags (symmetric memory when N > 1)
    nbytes = S.Group.block_bytes(N_total)
    join_used, hdl, mc = "single GPU: eight slabs into one y, no join", None, 0
    if world > 1 and args.join == "arrive":
        import torch.distributed._symmetric_memory as symm
        block = symm.empty(nbytes // 4, dtype=torch.float32, device=torch.device("cuda", local))
        block.zero_()
        hdl = symm.rendezvous(block, dist.group.WORLD)
        try:
            mc = int(hdl.multicast_ptr or 0)
        except Exception:
            mc = 0
        if os.environ.get("SPMV_NO_MULTICAST"):
            mc = 0
        join_used = "fused epilogue (%s stores into symmetric memory) + in-kernel arrival flags" % ("multicast" if mc else "peer")
    else:
        block = torch.zeros(nbytes // 4, dtype=torch.float32, device="cuda")
    use_group = world == 1 or args.join == "arrive"
    if use_group:
        grp = S.Group(C5_M, N_total, rank if world > 1 else 0, world if world > 1 else 1, block.data_ptr())
        if world > 1:
            grp.connect_ptrs([int(q) for q in hdl.buffer_ptrs], mc)
            torch.cuda.synchronize()
            dist.barrier()
        for g, p in zip(slab_ids, plans):
            grp.add(p, g * C5_SLAB_N)

        def y_view(ptr):
            off = (ptr - block.data_ptr()) // 4
            return block[off: off + N_total]

        def step(i, cs):
            grp.run(dx, cs)
        step_y = lambda: y_view(grp.run(dx, stream.cuda_stream))       # noqa: E731
    else:                                                 # N > 1 with the NCCL join: local slabs, then all_gather
        y_local = torch.zeros(len(slab_ids) * C5_SLAB_N, dtype=torch.float32, device="cuda")
        y_all = torch.zeros(N_total, dtype=torch.float32, device="cuda")
        join_used = "nccl all_gather_into_tensor"

        def step(i, cs):
            for k, p in enumerate(plans):
                p.run(dx, y_local[k * C5_SLAB_N:].data_ptr(), cs)
            dist.all_gather_into_tensor(y_all, y_local)

        def step_y():
            with torch.cuda.stream(stream):
                step(0, stream.cuda_stream)
            return y_all

    # ---- parity before timing: own slabs against the oracle, every rank's copy of every slab bit-identical ----
    with torch.cuda.stream(stream):
        y_dev = step_y()
        stream.synchronize()
        y_full = y_dev.cpu().numpy().copy()
        y_dev2 = step_y()                                  # the other y buffer, and run-to-run reproducibility
        stream.synchronize()
        y_full2 = y_dev2.cpu().numpy().copy()
    if use_group:
        grp.status()
    par = {}
    for k, g in enumerate(slab_ids):
        ys = y_full[g * C5_SLAB_N:(g + 1) * C5_SLAB_N]
        par[f"slab{g}_vs_oracle"] = record_parity(f"config 5 slab {g}", lambda: check_sampled(ys, samples[k], x, f"config 5 slab {g}"))
    par["two_calls_bit_identical"] = record_parity("reproducibility", lambda: (_ for _ in ()).throw(AssertionError("y differs between two calls"))
                                                   if y_full.tobytes() != y_full2.tobytes() else None)
    digs = [digest(y_full[g * C5_SLAB_N:(g + 1) * C5_SLAB_N]) for g in range(C5_SLABS)]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {"digs": digs, "par": par, "failed": list(PARITY_FAILED)})
        owner = {g: r for r in range(world) for g in range(r * C5_SLABS // world, (r + 1) * C5_SLABS // world)}
        bad = [f"slab {g}: rank {r} holds {gathered[r]['digs'][g]}, owner rank {owner[g]} computed {gathered[owner[g]]['digs'][g]}"
               for g in range(C5_SLABS) for r in range(world) if gathered[r]["digs"][g] != gathered[owner[g]]["digs"][g]]
        all_par = {}
        for r in range(world):
            all_par.update(gathered[r]["par"])
            for f in gathered[r]["failed"]:
                if f not in PARITY_FAILED:
                    PARITY_FAILED.append(f)
        all_par["gathered_y_identical_on_all_ranks"] = "ok" if not bad else "FAILED: " + "; ".join(bad[:3])
        if bad:
            PARITY_FAILED.extend(bad)
        par = all_par
    par["oracle"] = (f"oracle/spmv_oracle.c orc_csc_gemv (csr_naive.cu:14-22 semantics, fp32 sequential) + fp64, {C5_CHECK_COLS} seeded "
                     "columns per slab, tests/parity.py gate (1e-5 * sum|x a| per element, 1e-5 * max|y|, abs 1e-3)")

    # ---- the timed region: K steps ----------------------------------------------------------------
    ms, per = timed_steps(torch, step, args.steps, args.warmup, stream, graph=GRAPH, barrier=barrier)
    if use_group:
        grp.status()
    t = torch.tensor([ms, alg_rank, phys_rank, float(touched_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
    alg_all = c5_alg_bytes(float(t[3]))
    us_step = ms * 1e3 / args.steps
    value = alg_all / (us_step * 1e-6) / 1e9

    # the kernel alone (no join), this rank's slabs: the roofline's launch duration
    n_local = len(plans)
    dy_scratch = torch.zeros(C5_SLAB_N, dtype=torch.float32, device="cuda")
    k_steps = max(8, args.steps // 2)
    ms_k, per_k = timed_steps(torch, lambda i, cs: plans[i % n_local].run(dx, dy_scratch, cs), k_steps * n_local, 4, stream, graph=GRAPH)
    us_call = ms_k * 1e3 / (k_steps * n_local)

    # ---- end to end: pinned host x -> device, the step, this rank's columns of y -> pinned host ----
    e2e = None
    if use_group:
        hx = torch.from_numpy(x).pin_memory()
        own0, own_n = slab_ids[0] * C5_SLAB_N, len(slab_ids) * C5_SLAB_N
        hy = torch.empty(own_n, dtype=torch.float32).pin_memory()
        e_steps = max(4, args.steps // 2)
        e_steps += e_steps % 2
        for it in range(2 + e_steps):
            if it == 2:
                torch.cuda.synchronize()
                if barrier:
                    barrier()
                t0 = time.perf_counter()
            grp.run_host(hx.data_ptr(), hy.data_ptr(), own0, own_n)
        e2e_s = (time.perf_counter() - t0) / e_steps
        e2e_ok = record_parity("e2e path", lambda: (_ for _ in ()).throw(AssertionError("host-buffer path and device path disagree"))
                               if hy.numpy().tobytes() != y_full[own0: own0 + own_n].tobytes() else None)
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": round(alg_all / e2e_s / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": world * C5_M * 4,
               "d2h_bytes_per_step": N_total * 4, "us_per_step": round(e2e_s * 1e6, 2),
               "timer": "host perf_counter around spmv_mg_run_host (H2D x, the step, D2H of the rank's own columns of y, "
                        "synchronise), max over ranks; byte counts are summed over the ranks",
               "matches_device_path": e2e_ok}

    launches = args.steps * sum(p.info()["kernels_per_run"] for p in plans) + (args.steps if world > 1 and use_group else 0)
    info0 = plans[0].info()
    cfg = c5_config()
    if args.chunk_mode != C5_CHUNK_MODE:
        cfg["chunk_mode"] = args.chunk_mode
    extra["timing"] = {"resident": f"{len(plans)} slabs of {round(info0['device_bytes'] / 1e6, 1)} MB on this rank",
                       "launch": ("CUDA graphs of up to %d steps, %d timed chunks" % (GRAPH_STEPS, len(per))) if GRAPH else "one stream launch per call"}
    form = {4: "strips_kernel (+ strips_reduce_kernel)", 3: "panel_kernel<16,false,false,4,true> (lane-owned blocks)"}.get(args.chunk_mode, "panel_kernel (multi-row)")
    alg_call, phys_call = alg_rank / n_local, phys_rank / n_local
    join = {"used": join_used, "us_per_step": round(us_step, 3), "us_kernels_only_this_rank": round(us_call * n_local, 3),
            "us_join_overhead": round(us_step - us_call * n_local, 3)}
    extra["join"] = join
    extra["parity"] = par
    extra["build"] = {"slabs_this_rank": slab_ids, "generate_s": round(gen_s, 1), "pack_s": round(pack_s, 1), "nnz_this_rank": int(sum(nnz))}
    extra["us_per_step_stats"] = stats(per)
    extra["us_per_slab_call_stats"] = stats(per_k)
    roof = {"bound": "hbm", "achieved": round(alg_call / (us_call * 1e-6) / 1e9, 1), "peak": peak, "unit": "GB/s",
            "frac": round(alg_call / (us_call * 1e-6) / 1e9 / peak, 4), "traffic": ncu_traffic(f"c5/{HEADLINE}/mode{args.chunk_mode}"),
            "peak_source": peak_src, "kernel": form, "us_per_launch": round(us_call, 3),
            "alg_bytes_per_launch": alg_call, "phys_bytes_per_launch": phys_call,
            "phys_frac": round(phys_call / (us_call * 1e-6) / 1e9 / peak, 4),
            "launch": "one slab call (65536x131072): the strips kernel and its reduce kernel, timed alone on this rank"}

    # ---------------- N = 1 extras: every other config, per variant, each checked ------------------
    if world == 1 and not args.quick:
        for p in plans:
            p.close()
        plans.clear()
        if use_group:
            grp.close()
        del block
        torch.cuda.empty_cache()
        ob, parity = oracle_bindings()
        v_steps, v_warm = max(50, args.steps * 5), 10
        cfgs = {}
        for name in ("c2", "c1", "c3", "c0"):
            Mc, Nc, sac, sxc = synth.CONFIGS[name]
            Ac = synth.gen_matrix(Mc, Nc, sac)
            xc = synth.gen_vector(Mc, sxc)
            y32 = ob.sgemv_dense(Ac, xc)
            y64, sabs = ob.sgemv_dense_f64(Ac, xc)
            cfgs[name] = {"M": Mc, "N": Nc, "weight_sparsity": sac, "activation_sparsity": sxc,
                          "oracle": "orc_sgemv_dense (tester.cpp:36-45 restated), all columns"}
            for v in ("wsp", "asp", "awsp", "tcsr"):
                r, pl, _, _ = measure_variant(torch, S, v, lambda vv: S.Plan.from_dense(vv, Ac), xc, v_steps, v_warm, stream,
                                              check=lambda vv, y: parity.check_y(y, y32, y64, sabs, f"{name} {vv}"), hot=True,
                                              e2e=(name == "c2"))
                keep = ("us_per_call", "us_min", "us_median", "hot_l2_us_per_call", "e2e_us_per_call", "e2e_bytes", "alg_MB", "phys_MB",
                        "eff_GBps", "phys_GBps", "l2_copies", "parity", "kernels_per_call")
                cfgs[name][v] = {k: r[k] for k in keep if k in r}
                cfgs[name][v]["frac_alg"] = round(r["eff_GBps"] / peak, 4)
                cfgs[name][v]["frac_phys"] = round(r["phys_GBps"] / peak, 4)
                for p in pl:
                    p.close()
            if name == "c0" and not args.no_aux:
                extra["gpu_comparator"] = dict(gpu_comparators(torch, Ac, xc, y32, stream),
                                               shape="4096x4096, 50%/50% (test/main.cpp), the reference launchers' own cold-launch timing")
            if name == "c2" and not args.no_aux:
                cscp = ob.dense_to_csc(Ac)
                extra["cpu_baseline_c2"] = cpu_baselines(Ac, xc, (Nc, *cscp), step_alg_bytes(Ac, xc, ["awsp"]))
        # config 4: power-law row lengths, 1M x 1M, wsp (32-bit row ids, x gathered through L2)
        try:
            cp4, ri4, va4 = synth.powerlaw_csc(1 << 20, 1 << 20, seed=42)
            x4 = synth.gen_vector(1 << 20, 0.0, seed=7)
            rng = np.random.default_rng(404)
            cols4 = np.sort(rng.choice(1 << 20, size=4096, replace=False))
            lens = (cp4[cols4 + 1] - cp4[cols4]).astype(np.int64)
            scp = np.zeros(cols4.size + 1, np.int64)
            np.cumsum(lens, out=scp[1:])
            take = np.concatenate([np.arange(cp4[c], cp4[c + 1]) for c in cols4])
            sample4 = (cols4, scp, ri4[take].copy(), va4[take].copy())
            r, pl, _, _ = measure_variant(torch, S, "wsp", lambda vv: S.Plan.from_csc(vv, 1 << 20, 1 << 20, cp4, ri4, va4), x4,
                                          max(50, args.steps), 5, stream,
                                          check=lambda vv, y: check_sampled(y, sample4, x4, "config 4 wsp"))
            cfgs["c4"] = {"M": 1 << 20, "N": 1 << 20, "workload": "power-law row lengths (Pareto alpha=2, ~16 nnz/row), dense x",
                          "oracle": "orc_csc_gemv on 4096 seeded columns", "wsp": dict(r, frac_alg=round(r["eff_GBps"] / peak, 4))}
            for p in pl:
                p.close()
        except Exception as e:                            # never lose the headline line to an auxiliary config
            cfgs["c4"] = {"error": str(e)[:200]}
        extra["configs"] = cfgs

    # ---------------- the CPU baseline beside the headline (rank 0, N = 1) -------------------------
    if world == 1 and not args.quick and not args.no_aux:
        cores = os.cpu_count() or 1
        cp, ri, va = c5_slab_csc(synth, 0)
        ob, _ = oracle_bindings()
        alg_slab = alg_all / C5_SLABS
        ob.csc_gemv(C5_SLAB_N, cp, ri, va, x, threads=cores)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ob.csc_gemv(C5_SLAB_N, cp, ri, va, x, threads=cores)
        dt_all = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        ob.csc_gemv(C5_SLAB_N, cp, ri, va, x, threads=1)
        dt_1 = time.perf_counter() - t0
        extra["cpu_baseline"] = {"value": round(alg_slab / dt_all / 1e9, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"one of the 8 slabs (65536x131072, {ri.size} non-zeros) through the oracle's CSR(A^T) restatement of "
                                           f"csr_naive.cu:14-22, OpenMP on {cores} host threads, best-effort mean of {reps}: {dt_all * 1e3:.1f} ms per slab",
                                 "csr_1t": {"GBps": round(alg_slab / dt_1 / 1e9, 4), "ms_per_slab": round(dt_1 * 1e3, 1)},
                                 "csr_all_cores": {"GBps": round(alg_slab / dt_all / 1e9, 4), "ms_per_slab": round(dt_all * 1e3, 1), "threads": cores},
                                 "nproc": cores}

    clocks = sampler.result()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 6), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "us_per_step": round(us_step, 3), "roofline": roof,
                "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                    "note": "the NCCL-join comparison run has no host-buffer path"},
                "gpu_launches": launches, "clocks": clocks, "parity_ok": not PARITY_FAILED}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if PARITY_FAILED:
        print("PARITY FAILED:\n  " + "\n  ".join(PARITY_FAILED[:10]), file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — sparse SGEMV Y = x·A on B200: µs/call and effective HBM GB/s vs the roofline.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                    (the reference's CPU path, rank 0)

N = 1   workload = BASELINE config 2 (LLM decode FFN up-proj: A 4096x14336, 70 % weight-sparse,
        x 50 % activation-sparse).  The config names three variants (wsp; asp/awsp), so a step is
        one call of each of them on the same A and x; every variant (and tcsr, and the other
        single-GPU configs) is also timed alone and reported under "variants" / "configs", and
        "roofline" describes the step's dominant (longest) kernel.
N > 1   workload = BASELINE config 5 family, weak scaling: every rank owns a 131072-column slab
        of A (65536 rows, 99 % sparse, built directly in sparse form; lane-owned block form of the
        awsp format, chunk_mode 3), x is replicated, a step is the local awsp call whose epilogue
        stores its slice of Y into every rank's buffer (fused all-gather; the NCCL all-gather join
        is timed beside it, "join").  At N = 8 this is exactly config 5 (65536 x 1048576).  The N = 1 line also carries this slab's single-GPU number
        ("weak_scaling_unit") so per-N efficiency can be computed on one workload.

value   = algorithmic bytes (8*nnz_touched + 4(N+1) + 4M + 4N) of all ranks divided by
          the device time of the timed region (CUDA events, max over ranks), in GB/s.
L2      every timed loop rotates over clones of the packed matrix whose total size exceeds
        2.5x the 126 MB L2, so each call streams from HBM (config.l2 says how many copies).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sparse SGEMV Y=xA effective HBM throughput (algorithmic bytes / device time)"
UNIT = "GB/s"
L2_BYTES = 126e6
HEADLINE = "awsp"                      # multi-GPU (config 5) variant
STEP_VARIANTS = ("wsp", "asp", "awsp")   # config 2 names all three: one step = one call of each
C5_M, C5_SLAB_N, C5_DENSITY, C5_SX = 65536, 131072, 0.01, 0.5


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
def make_copies(plan, want_bytes=2.5 * L2_BYTES, max_copies=12):
    """Clones of the resident matrix so that a rotation over them defeats the L2."""
    info = plan.info()
    each = max(1, info["device_bytes"])
    n = int(min(max_copies, max(1, -(-want_bytes // each))))
    return [plan] + [plan.clone() for _ in range(n - 1)]


GRAPH_STEPS = 100   # steps captured per CUDA graph
GRAPH = True        # --no-graph clears it


def timed_steps(torch, step, steps, warmup, stream, graph=True):
    """Device time (ms) of exactly `steps` calls of step(i, cuda_stream) on `stream`, CUDA events,
    device-wide synchronize on both sides.  With graph=True the steps are captured into CUDA
    graphs (GRAPH_STEPS per graph, replayed back to back; a second graph for the remainder):
    kernels launched one by one into a stream start on a ~2 us dispatch cadence on this driver,
    a graph runs them back to back (tools/graph_vs_stream.py: 2.0-2.6 us per call)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for i in range(warmup):
            step(i, stream.cuda_stream)
        stream.synchronize()
        if not graph:
            torch.cuda.synchronize()
            e0.record(stream)
            for i in range(steps):
                step(i, stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        per = min(GRAPH_STEPS, steps)
        reps, rem = divmod(steps, per)
        graphs = []
        for n in ([per] if rem == 0 else [per, rem]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                cs = torch.cuda.current_stream().cuda_stream
                for i in range(n):
                    step(i, cs)
            graphs.append(g)
        graphs[0].replay()                                 # untimed: first replay uploads the graph
        if rem:
            graphs[1].replay()
        stream.synchronize()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            graphs[0].replay()
        if rem:
            graphs[1].replay()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def time_loop(torch, plans, dx, dy, steps, warmup, stream, graph=True):
    """`steps` calls rotating over `plans` (clones of one matrix, > 2.5x L2 in total)."""
    n = len(plans)
    return timed_steps(torch, lambda i, cs: plans[i % n].run(dx, dy, cs), steps, warmup, stream, graph)


def measure_variant(torch, S, variant, build, x, steps, warmup, stream, check=None):
    """Pack, clone, time.  Returns a dict; `build(variant)` returns a Plan."""
    t0 = time.perf_counter()
    plan = build(variant)
    pack_s = time.perf_counter() - t0
    info = plan.info()
    alg, phys, nnz_t = plan.traffic(x)
    plans = make_copies(plan)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(info["N"], dtype=torch.float32, device="cuda")
    if check is not None:
        plan.run(dx, dy, stream.cuda_stream)
        stream.synchronize()
        check(variant, dy.cpu().numpy())
    ms = time_loop(torch, plans, dx, dy, steps, warmup, stream, graph=GRAPH)
    us = ms * 1e3 / steps
    res = {"us_per_call": round(us, 3), "alg_MB": round(alg / 1e6, 3), "phys_MB": round(phys / 1e6, 3),
           "eff_GBps": round(alg / (us * 1e-6) / 1e9, 1), "phys_GBps": round(phys / (us * 1e-6) / 1e9, 1),
           "nnz_touched": nnz_t, "l2_copies": len(plans), "resident_MB": round(info["device_bytes"] / 1e6, 1),
           "grid": [info["grid_x"], info["grid_y"]], "kernels_per_call": info["kernels_per_run"],
           "slab_cols": info["slab_cols"], "row_splits": info["row_splits"], "pack_s": round(pack_s, 2)}
    return res, plans, (dx, dy), (alg, phys)


def e2e_loop(plan, x, N, steps, warmup, torch):
    """The reference launcher's per-call part through the C-ABI with HOST buffers:
    H2D x, kernels, D2H y, synchronise — every step (spmv_run_host)."""
    hx = torch.from_numpy(x).pin_memory()
    hy = torch.empty(N, dtype=torch.float32).pin_memory()
    for _ in range(warmup):
        plan.run_host_ptr(hx.data_ptr(), hy.data_ptr())
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.run_host_ptr(hx.data_ptr(), hy.data_ptr())
    dt = time.perf_counter() - t0
    return dt / steps, hx.numel() * 4, hy.numel() * 4, hy.numpy().copy()


# ------------------------------------------------------------------------------------------------
def step_alg_bytes(A, x, names):
    """Algorithmic bytes of one step = one call of each variant in `names`."""
    M, N = A.shape
    nnz = int(np.count_nonzero(A))
    act = x != 0
    nnz_t = int(np.count_nonzero(A[act]))
    mnz = int(np.count_nonzero(act))
    vec = 4.0 * M + 4.0 * N
    per = {"wsp": 8.0 * nnz + 4 * (N + 1) + vec, "tcsr": 8.0 * nnz_t + 4 * (N + 1) + vec,
           "awsp": 8.0 * nnz_t + 4 * (N + 1) + vec, "asp": 4.0 * mnz * N + vec}
    return sum(per[v] for v in names)


def cpu_reference_sampled(A, x, alg_bytes, steps, warmup, calls_per_step=1, budget_s=100.0):
    """The reference's own CPU path (SgemvCPU, tester.cpp:36-45) from oracle/_ref when it was
    built, else the oracle's restatement of it.  The reference has one CPU implementation for
    every variant, so a step of `calls_per_step` SGEMV calls is that many SgemvCPU calls.  Each
    call runs on the first `rows` rows of the full-width matrix (row stride stays N, as in the
    reference), `rows` chosen so the run fits the time budget; throughput is scaled by rows/M."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bindings as ob
    if ob.have_ref_cpu():
        fn, kind = ob.ref_sgemv_cpu, "reference"
    else:
        fn, kind = ob.sgemv_dense, "port"
    M, N = A.shape
    probe_rows = 64
    t0 = time.perf_counter()
    fn(A[:probe_rows], x[:probe_rows])
    per_row = (time.perf_counter() - t0) / probe_rows
    rows = int(min(M, max(32, budget_s / max(1, (steps + warmup) * calls_per_step) / per_row)))
    rows = max(32, rows // 32 * 32)
    Ar, xr = np.ascontiguousarray(A[:rows]), np.ascontiguousarray(x[:rows])
    for _ in range(warmup * calls_per_step):
        fn(Ar, xr)
    t0 = time.perf_counter()
    for _ in range(steps * calls_per_step):
        fn(Ar, xr)
    dt = (time.perf_counter() - t0) / steps
    frac = rows / M
    gbps = alg_bytes * frac / dt / 1e9
    what = "the reference SgemvCPU (tester.cpp:36-45, oracle/_ref)" if kind == "reference" else "the oracle port of SgemvCPU"
    sample = (f"{steps} steps of {calls_per_step} call(s) of {what} on the first {rows} of {M} rows of the "
              f"full-width {M}x{N} matrix (dense loop, 1 thread, {dt / calls_per_step * 1e3:.1f} ms per call; "
              f"full-matrix call ~{dt / calls_per_step / frac * 1e3:.0f} ms)")
    return gbps, dt, kind, sample, frac


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from spmv_test_b200 import synth
    if args.gpus > 1:
        # this arm's N > 1 workload is the config-5 family; the reference's CPU path is the dense loop,
        # and a slab cannot exist in dense form (34 GB), so the bounded sample is a dense block of
        # the first 1024 columns of one slab (outputs are independent: time is linear in columns)
        cols = 1024
        cp, ri, va = synth.bernoulli_csc(C5_M, cols, C5_DENSITY, seed=5000)
        A = np.zeros((C5_M, cols), np.float32)
        A[ri, np.repeat(np.arange(cols), np.diff(cp))] = va
        x = synth.gen_vector(C5_M, C5_SX, seed=4321)
        alg = 8.0 * float(np.count_nonzero(A[x != 0.0])) + 4.0 * (cols + 1) + 4.0 * C5_M + 4.0 * cols
        names = [HEADLINE]
        cfg = workload_config("c5", HEADLINE, None)
        cfg["N_total"] = args.gpus * C5_SLAB_N
        note = (f"config-5 slab sampled as a dense {C5_M}x{cols} block of its first columns (the job is "
                f"{args.gpus * C5_SLAB_N // cols} such blocks, processed one after the other on one thread); ")
    else:
        M, N, sa, sx = synth.CONFIGS["c2"]
        A = synth.gen_matrix(M, N, sa)
        x = synth.gen_vector(M, sx)
        names = list(STEP_VARIANTS)
        alg = step_alg_bytes(A, x, names)
        cfg = workload_config("c2", "+".join(names), None)
        note = ""
    gbps, dt, kind, sample, frac = cpu_reference_sampled(A, x, alg, args.steps, args.warmup, calls_per_step=len(names))
    line = {"impl": "reference", "metric": METRIC, "value": round(gbps, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": round(gbps, 4), "unit": UNIT, "cores": 1, "kind": kind, "sample": note + sample},
            "e2e": {"value": round(gbps, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(cfg, variant, l2):
    if cfg == "c2":
        return {"workload": "BASELINE config 2: A 4096x14336 fp32, 70% weight-sparse, x 50% activation-sparse; one step = "
                            f"one SGEMV call of each variant in [{variant}] on the same A and x, seeds 1234/4321",
                "M": 4096, "N": 14336, "weight_sparsity": 0.7, "activation_sparsity": 0.5, "variant": variant,
                "l2": l2}
    return {"workload": "BASELINE config 5 family (weak scaling): per GPU a 131072-column slab of A "
                        "(65536 rows, 99% sparse, built in sparse form), x 50% activation-sparse, awsp (lane-owned "
                        "blocks, chunk_mode 3) + all-gather of Y; 8 GPUs = 65536x1048576",
            "M": C5_M, "N_per_gpu": C5_SLAB_N, "weight_sparsity": 0.99, "activation_sparsity": C5_SX,
            "variant": variant, "l2": l2}


# ------------------------------------------------------------------------------------------------
# Config 5 (1 % dense, half of the activations non-zero) runs on the lane-owned block form of the awsp
# format (chunk_mode 3: every stored non-zero is read, x is a multiplier, one pass per chunk); the
# row-addressable multi-row form (chunk_mode 0) is reported beside it on one GPU.
C5_CHUNK_MODE = 3


def slab_unit(torch, S, synth, rank, steps, warmup, stream, chunk_mode=C5_CHUNK_MODE):
    """One GPU's config-5 slab: build, time the kernel alone.  Returns (res, plans, bufs, x)."""
    col_ptr, row_idx, vals = synth.bernoulli_csc(C5_M, C5_SLAB_N, C5_DENSITY, seed=5000 + rank)
    x = synth.gen_vector(C5_M, C5_SX, seed=4321)

    def build(v):
        return S.Plan.from_csc(v, C5_M, C5_SLAB_N, col_ptr, row_idx, vals, chunk_mode=chunk_mode)
    res, plans, bufs, bytes_ = measure_variant(torch, S, HEADLINE, build, x, steps, warmup, stream)
    res["chunk_mode"] = chunk_mode
    return res, plans, bufs, x, bytes_


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline variant only (used under ncu)")
    ap.add_argument("--no-aux", action="store_true", help="skip config 4 / config-5 slab / CPU baseline legs")
    ap.add_argument("--no-graph", action="store_true", help="launch every call into the stream instead of replaying CUDA graphs")
    ap.add_argument("--variant", default=None, help="headline variant override (for profiling one kernel)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 10 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_reference_arm(args)
        return
    global GRAPH
    GRAPH = not args.no_graph
    args.steps = 2000 if args.steps is None else args.steps
    args.warmup = 50 if args.warmup is None else max(3, args.warmup)

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
    stream = torch.cuda.Stream()
    peak, peak_src = peaks()
    sampler = ClockSampler(local)
    sampler.start()
    extra = {}

    if world == 1:
        # ---------------- single GPU: config 2, one step = wsp + asp + awsp -----------------------
        M, N, sa, sx = synth.CONFIGS["c2"]
        A = synth.gen_matrix(M, N, sa)
        x = synth.gen_vector(M, sx)
        names = [args.variant] if args.variant else list(STEP_VARIANTS)
        variants, sets, algs, physs = {}, {}, {}, {}
        for v in names:
            res, plans, (dx, dy), (alg, phys) = measure_variant(
                torch, S, v, lambda vv: S.Plan.from_dense(vv, A), x, args.steps, args.warmup, stream)
            variants[v], sets[v], algs[v], physs[v] = res, plans, alg, phys
        # the timed region: K steps, each one call of every variant the config names
        def step(i, cs):
            for v in names:
                pl = sets[v]
                pl[i % len(pl)].run(dx, dy, cs)
        ms = timed_steps(torch, step, args.steps, args.warmup, stream, graph=not args.no_graph)
        alg_step = sum(algs.values())
        value = alg_step / (ms / args.steps * 1e-3) / 1e9
        # end to end through the host-buffer C-ABI call (H2D x, kernels, D2H y, synchronise)
        e2e_s, h2d, d2h = 0.0, 0, 0
        for v in names:
            plans = sets[v]
            plans[0].run(dx, dy, stream.cuda_stream)
            stream.synchronize()
            y_dev = dy.cpu().numpy()
            t, bi, bo, y_e2e = e2e_loop(plans[0], x, N, args.steps, args.warmup, torch)
            assert y_e2e.tobytes() == y_dev.tobytes(), f"{v}: host-buffer path and device path disagree"
            variants[v]["e2e_us_per_call"] = round(t * 1e6, 2)
            e2e_s += t
            h2d += bi
            d2h += bo
        launches = args.steps * sum(sets[v][0].info()["kernels_per_run"] for v in names)
        l2 = ("rotation over resident copies of every packed matrix (" +
              ", ".join(f"{v}: {len(sets[v])} x {variants[v]['resident_MB']} MB" for v in names) + "), each > 2.5x L2 in total")
        dominant = max(names, key=lambda v: variants[v]["us_per_call"])
        for v in names:
            for p in sets[v][1:]:
                p.close()
        if not args.quick:
            for v in [v for v in ("wsp", "asp", "awsp", "tcsr") if v not in names]:
                r, pl, _, _ = measure_variant(torch, S, v, lambda vv: S.Plan.from_dense(vv, A), x,
                                              args.steps, args.warmup, stream)
                variants[v] = r
                for p in pl:
                    p.close()
            cfgs = {}
            for name in ("c1", "c3", "c0"):
                Mc, Nc, sac, sxc = synth.CONFIGS[name]
                Ac = synth.gen_matrix(Mc, Nc, sac)
                xc = synth.gen_vector(Mc, sxc)
                cfgs[name] = {"M": Mc, "N": Nc, "weight_sparsity": sac, "activation_sparsity": sxc}
                for v in ("wsp", "asp", "awsp", "tcsr"):
                    r, pl, _, _ = measure_variant(torch, S, v, lambda vv: S.Plan.from_dense(vv, Ac), xc,
                                                  args.steps, args.warmup, stream)
                    cfgs[name][v] = {k: r[k] for k in ("us_per_call", "alg_MB", "phys_MB", "eff_GBps", "phys_GBps")}
                    for p in pl:
                        p.close()
            extra["configs"] = cfgs
        if not args.quick and not args.no_aux:
            # batched (multi-vector) wsp: A streamed once for 4 activation vectors
            try:
                bp = [S.Plan.from_dense("wsp", A)]
                bp += [bp[0].clone() for _ in range(2)]
                Xb = np.stack([synth.gen_vector(M, sx, seed=10 + b) for b in range(4)])
                dXb = torch.from_numpy(Xb).cuda()
                dYb = torch.zeros((4, N), device="cuda")
                alg_b = sum(bp[0].traffic(Xb[b])[0] for b in range(4))
                nb = max(50, args.steps // 10)
                ms_b = timed_steps(torch, lambda i, cs: bp[i % 3].run_batch(dXb, dYb, cs), nb, 5, stream, graph=GRAPH)
                us_b = ms_b * 1e3 / nb
                extra["batched_wsp"] = {"batch": 4, "us_per_batched_call": round(us_b, 3), "us_per_vector": round(us_b / 4, 3),
                                        "eff_GBps": round(alg_b / (us_b * 1e-6) / 1e9, 1),
                                        "note": "A's bytes are read once per 4 vectors, so this exceeds the single-vector HBM roofline"}
                for p in bp:
                    p.close()
            except Exception as e:
                extra["batched_wsp"] = {"error": str(e)[:200]}
            r, pl, _, _, _ = slab_unit(torch, S, synth, 0, max(50, args.steps // 10), 5, stream)
            extra["weak_scaling_unit"] = dict(r, workload="one GPU's config-5 slab (65536x131072, 99% sparse, x 50%), kernel only, "
                                                          "lane-owned blocks (chunk_mode 3)")
            for p in pl:
                p.close()
            r, pl, _, _, _ = slab_unit(torch, S, synth, 0, max(50, args.steps // 10), 5, stream, chunk_mode=0)
            extra["weak_scaling_unit_row_form"] = dict(r, workload="the same slab in the row-addressable multi-row form (chunk_mode 0)")
            for p in pl:
                p.close()
            # config 4: power-law row lengths, 1M x 1M, wsp (32-bit row ids, x gathered through L2)
            try:
                cp, ri, va = synth.powerlaw_csc(1 << 20, 1 << 20, seed=42)
                x4 = synth.gen_vector(1 << 20, 0.0, seed=7)
                r, pl, _, _ = measure_variant(torch, S, "wsp", lambda vv: S.Plan.from_csc(vv, 1 << 20, 1 << 20, cp, ri, va),
                                              x4, max(50, args.steps // 10), 5, stream)
                extra["config4_powerlaw"] = dict(r, workload="1Mx1M power-law rows (Pareto alpha=2, ~16 nnz/row), wsp, dense x")
                for p in pl:
                    p.close()
            except Exception as e:  # never lose the headline line to an auxiliary config
                extra["config4_powerlaw"] = {"error": str(e)[:200]}
            cpu_gbps, cpu_dt, kind, sample, _ = cpu_reference_sampled(A, x, alg_step, 2, 1, calls_per_step=len(names),
                                                                      budget_s=20.0)
            extra["cpu_baseline"] = {"value": round(cpu_gbps, 4), "unit": UNIT, "cores": 1, "kind": kind, "sample": sample}
        extra["variants"] = variants
        extra["step"] = names
        cfg = workload_config("c2", "+".join(names), l2)
        cfg["launch"] = ("CUDA graphs of %d steps replayed back to back" % min(GRAPH_STEPS, args.steps)) if GRAPH else "one stream launch per call"
        roof_alg, roof_us, phys = algs[dominant], variants[dominant]["us_per_call"], physs[dominant]
        roof_kernel = {"wsp": "wsp_ring_kernel<uint2>", "asp": "asp_kernel", "awsp": "panel_kernel<16,false,false,8>",
                       "tcsr": "panel_kernel<16,true,false,8>"}[dominant]
        e2e_val = alg_step / e2e_s / 1e9
        scaling = "weak"
        traffic = ncu_traffic(f"c2/{dominant}")
    else:
        # ---------------- multi GPU: config-5 slabs + all-gather ------------------------------
        res, plans, (dx, dy), x, (alg, phys) = slab_unit(torch, S, synth, rank, max(20, args.steps // 10), 5, stream)
        bounds = [g * C5_SLAB_N for g in range(world + 1)]
        n = len(plans)
        state = {"i": 0}

        def local_run(d_x, d_y):
            plans[state["i"] % n].run(d_x, d_y, torch.cuda.current_stream().cuda_stream)
            state["i"] += 1
        class _Rot:                                       # the sharded runner sees a rotating plan
            def run(self, d_x, d_y, stream=None):
                local_run(d_x, d_y)

            def run_scatter(self, d_x, ptrs, offset, mc=0, stream=None):
                plans[state["i"] % n].run_scatter(d_x, ptrs, offset, mc, torch.cuda.current_stream().cuda_stream)
                state["i"] += 1
        rot = _Rot()

        def timed_join(join):
            sh = S.ShardedSgemv(bounds, rank, world, plan=rot, join=join)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                for _ in range(args.warmup):
                    y_full = sh.run(dx)
                stream.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(args.steps):
                    y_full = sh.run(dx)
                e1.record(stream)
                stream.synchronize()
                torch.cuda.synchronize()
                dist.barrier()
            return sh, y_full, e0.elapsed_time(e1)

        sh_nccl, y_nccl, ms_nccl = timed_join("nccl")
        try:
            sh, y_full, ms_fused = timed_join("fused")
            assert torch.equal(y_full, y_nccl), "fused epilogue and NCCL all-gather disagree"
            join_used = "fused epilogue (%s stores into symmetric memory + barrier)" % ("multicast" if sh.multicast else "peer")
        except Exception as e:                            # no symmetric memory on this box: NCCL join
            sh, y_full, ms_fused = sh_nccl, y_nccl, None
            join_used = "nccl all_gather (fused epilogue unavailable: %s)" % str(e)[:120]
        ms_join = ms_fused if ms_fused is not None else ms_nccl
        e_ms = torch.tensor([ms_join], dtype=torch.float64, device="cuda")
        t = torch.tensor([ms_join, alg, phys], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        alg_all = float(t[1])
        value = alg_all / (ms / args.steps * 1e-3) / 1e9
        # end to end: pinned host x -> device, local kernel, all-gather, full y -> pinned host
        hx = torch.from_numpy(x).pin_memory()
        hy = torch.empty(world * C5_SLAB_N, dtype=torch.float32).pin_memory()
        e2e_steps = max(10, args.steps // 10)
        with torch.cuda.stream(stream):
            for it in range(3 + e2e_steps):
                if it == 3:
                    stream.synchronize()
                    dist.barrier()
                    t0 = time.perf_counter()
                dx.copy_(hx, non_blocking=True)
                y_full = sh.run(dx)
                hy.copy_(y_full, non_blocking=True)
                stream.synchronize()
            e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e_val = alg_all / e2e_s / 1e9
        h2d, d2h = hx.numel() * 4, hy.numel() * 4
        launches = args.steps * plans[0].info()["kernels_per_run"]
        l2 = (f"rotation over {len(plans)} resident copies of the slab ({res['resident_MB']} MB each): "
              "inputs larger than L2")
        cfg = workload_config("c5", HEADLINE, l2)
        cfg["N_total"] = world * C5_SLAB_N
        extra["variants"] = {HEADLINE + "_kernel_only_rank0": res}
        extra["allgather_bytes_per_step"] = world * C5_SLAB_N * 4
        tn = torch.tensor([ms_nccl], dtype=torch.float64, device="cuda")
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        extra["join"] = {"used": join_used, "us_per_step_fused": None if ms_fused is None else round(ms / args.steps * 1e3, 3),
                         "us_per_step_nccl": round(float(tn[0]) / args.steps * 1e3, 3),
                         "us_kernel_only_rank0": res["us_per_call"]}
        roof_alg, roof_us = alg, res["us_per_call"]
        roof_kernel = "panel_kernel<16,false,false,4,true> (lane-owned blocks)"
        scaling = "weak"
        traffic = ncu_traffic(f"c5/{HEADLINE}")

    clocks = sampler.result()
    if rank == 0:
        achieved = roof_alg / (roof_us * 1e-6) / 1e9
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 6), "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "us_per_call": round(ms / args.steps * 1e3, 3),
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                             "kernel": roof_kernel,
                             "alg_bytes_per_launch": roof_alg,
                             "phys_bytes_per_launch": phys, "phys_frac": round(phys / (roof_us * 1e-6) / 1e9 / peak, 4)},
                "e2e": {"value": round(e2e_val, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "us_per_call": round(e2e_s * 1e6, 2),
                        "timer": "host perf_counter around the synchronous host-buffer call, max over ranks"},
                "gpu_launches": launches, "clocks": clocks}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Parity gate shared by the GPU tests, smoke() and bench.py.

fp32 results of two valid summation orders cannot agree to 1e-5 relative to |y_i| (cancellation;
SURVEY §7 hard part 8), so "1e-5 relative" is normalised the way error bounds for dot products
are: per element by the magnitude sum  S_i = sum_j |x_j a_ji|, and in the max-norm by ||y_ref||.
"""
import numpy as np

REL_TOL = 1e-5          # north_star: within 1e-5 relative (fp32)
REF_ABS_TOL = 1e-3      # the reference's own CompareY threshold (tester.cpp:75)


def check_y(y, y_ref32, y_ref64, abs_sum, what=""):
    y = np.asarray(y, np.float64)
    assert np.all(np.isfinite(y)), f"{what}: non-finite output"
    tiny = 1e-30
    for name, ref in (("fp32 sequential oracle", np.asarray(y_ref32, np.float64)), ("fp64 oracle", y_ref64)):
        err = np.abs(y - ref)
        worst = float(np.max(err / (abs_sum + tiny))) if err.size else 0.0
        assert worst <= REL_TOL, f"{what}: |y-y_ref|/sum|x a| = {worst:.3e} > {REL_TOL} vs {name}"
        norm = float(np.max(np.abs(ref))) if ref.size else 0.0
        assert (float(np.max(err)) if err.size else 0.0) <= REL_TOL * norm + tiny, \
            f"{what}: max-norm error {np.max(err):.3e} > {REL_TOL}*{norm:.3e} vs {name}"
    n_bad = int(np.sum(np.abs(y - np.asarray(y_ref32, np.float64)) > REF_ABS_TOL))
    assert n_bad == 0, f"{what}: {n_bad} elements beyond the reference's abs 1e-3 gate"

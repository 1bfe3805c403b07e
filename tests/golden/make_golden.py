"""Generates tests/golden/ref_golden.json and ref_golden_y.npz from the REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference compiled in place into
oracle/_ref/libspmv_ref_cpu.so by oracle/Makefile; nothing of the reference is copied):

    make -C oracle && python tests/golden/make_golden.py

The reference ships no golden vectors, so these are minted from its own host code:
for each seeded case, SHA-256 digests of every stream the six reference packers produce
(CSRMatrix, TCSRMatrix, WSPMatrix, ASPMatrix, AWSPMatrix, AWSPRefMatrix) and the y computed by
SparseSgemvTester::SgemvCPU (tester.cpp:36-45), stored in full for the small cases.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_bindings as ob  # noqa: E402

# (name, M, N, weight sparsity, activation sparsity, seed)
CASES = [
    ("t32", 32, 32, 0.5, 0.5, 1),
    ("t64x32_dense", 64, 32, 0.0, 0.0, 2),
    ("t128x64", 128, 64, 0.9, 0.5, 3),
    ("t256x512", 256, 512, 0.7, 0.5, 4),
    ("t1024x128", 1024, 128, 0.5, 0.9, 5),      # M % 1024 == 0: the reference wsp kernel's own constraint
    ("t512x96_allzero_cols", 512, 96, 0.97, 0.3, 6),
    ("t2048x256", 2048, 256, 0.7, 0.5, 7),
]
LAYOUTS = ["csr", "tcsr", "wsp", "asp", "awsp", "awsp_ref"]


def digest(a):
    if a is None:
        return None
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() + f":{a.size}"


def case_inputs(M, N, sa, sx, seed):
    A = ob.gen_matrix(M, N, sa, seed)
    x = ob.gen_vector(M, sx, seed + 1000)
    if M >= 64:
        A[3, :] = 0.0          # an all-zero row
        A[:, 5] = 0.0          # an all-zero column
        A[7, 9] = -0.0         # negative zero is a zero for every packer (== 0.0f)
        x[11] = -0.0
    return A, x


def main():
    assert ob.have_ref_cpu(), "build oracle/_ref first (make -C oracle)"
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libspmv_ref_cpu.so (reference compiled in place)",
           "cases": {}}
    ys = {}
    for name, M, N, sa, sx, seed in CASES:
        A, x = case_inputs(M, N, sa, sx, seed)
        entry = {"M": M, "N": N, "sa": sa, "sx": sx, "seed": seed, "A": digest(A), "x": digest(x), "layouts": {}}
        for lay in LAYOUTS:
            p = ob.ref_pack(lay, A)
            entry["layouts"][lay] = {"i32_a": digest(p.i32_a), "i32_b": digest(p.i32_b), "u32": digest(p.u32),
                                     "f32": digest(p.f32), "aux": [int(v) for v in p.aux]}
        y = ob.ref_sgemv_cpu(A, x)
        entry["y"] = digest(y)
        ys[name] = y
        out["cases"][name] = entry
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "ref_golden_y.npz"), **ys)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()

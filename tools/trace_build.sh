#!/bin/bash
# Development tool: builds spmv_test_b200/lib/libspmv_b200_trace.so with -DSPMV_TRACE (per-warp
# %globaltimer stamps inside the kernels).  Never loaded by the package, tests or bench.
set -euo pipefail
cd "$(dirname "$0")/.."
CS=spmv_test_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -DSPMV_TRACE -rdc=true -Iinclude -I$CS \
     -Xcompiler -fPIC,-fvisibility=hidden -shared -o spmv_test_b200/lib/libspmv_b200_trace.so \
     $CS/capi.cu $CS/wsp.cu $CS/asp.cu $CS/panel.cu $CS/panel_rs.cu $CS/strips.cu $CS/mg.cu $CS/compact.cu $CS/pack_dev.cu $CS/pack_host.cpp -cudart static

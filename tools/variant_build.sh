#!/bin/bash
# Development tool: build spmv_test_b200/lib/libspmv_b200_<tag>.so from the working tree with extra
# compiler flags, for same-box A/B timing (select it with SPMV_B200_LIB=...):
#     tools/variant_build.sh <tag> [-DSPMV_STRIP_REGS=0 ...]
set -euo pipefail
cd "$(dirname "$0")/.."
TAG=$1; shift
CS=spmv_test_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a "$@" -Iinclude -I$CS \
     -Xcompiler -fPIC,-fvisibility=hidden -shared -o spmv_test_b200/lib/libspmv_b200_$TAG.so \
     $CS/capi.cu $CS/wsp.cu $CS/asp.cu $CS/panel.cu $CS/panel_rs.cu $CS/strips.cu $CS/mg.cu $CS/compact.cu $CS/pack_dev.cu $CS/pack_host.cpp -cudart static

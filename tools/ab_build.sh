#!/bin/bash
# Development tool: build an alternative library with one source file taken from another commit,
# for same-box A/B timing:  tools/ab_build.sh <commit> <file under spmv_test_b200/csrc> <tag>
set -euo pipefail
cd "$(dirname "$0")/.."
C=$1; F=$2; TAG=$3
D=$(mktemp -d)
cp spmv_test_b200/csrc/* "$D"/
git show "$C:spmv_test_b200/csrc/$F" > "$D/$F"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Iinclude -I"$D" \
     -Xcompiler -fPIC,-fvisibility=hidden -shared -o "spmv_test_b200/lib/libspmv_b200_$TAG.so" \
     "$D"/capi.cu "$D"/wsp.cu "$D"/asp.cu "$D"/panel.cu "$D"/compact.cu "$D"/pack_dev.cu "$D"/pack_host.cpp -cudart static
rm -rf "$D"

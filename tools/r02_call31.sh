#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
for rep in 1 2; do for lib in "" _aspr32 _aspr40 _aspr48; do for cfg in c2 c0 c3; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py asp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c31_asp.log
done; done; done
for lib in "" _aspr32 _aspr48; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 300 python tools/sweep.py asp c2 row_splits=4,5,6,8,10 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c31_asp.log
done

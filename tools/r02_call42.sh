#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
for cfg in c0 c3 c2; do
  timeout 300 python tools/sweep.py asp $cfg row_splits=0,8,12,16,18,24,32,37 2>&1 | tee -a $O/c42_asp.log
done
for lib in "" _st6 _st10; do for cfg in c2 c0 c3; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py wsp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c42_wsp.log
done; done

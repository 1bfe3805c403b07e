#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c21_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/c21_pytest.log
for rep in 1 2; do for lib in "" _wspring; do for cfg in c2 c0 c1 c3; do
  SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 200 python tools/sweep.py wsp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c21_wsp.log
done; done; done

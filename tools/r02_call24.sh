#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200_wspbulk.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small or config or edge or options or batch or panels or csc or relu" > $O/c24_pytest.log 2>&1; echo "pytest(bulk) rc=$?"
tail -4 $O/c24_pytest.log
for rep in 1 2; do for lib in "" _wspbulk; do for cfg in c2 c0 c1 c3; do
  SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 200 python tools/sweep.py wsp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c24_wsp.log
done; done; done
for lib in "" _wspbulk; do SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 300 python tools/batch_bench.py c2 wsp 2>&1 | sed "s/^/[lib$lib] /" | tail -4; done

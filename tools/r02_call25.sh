#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
for v in awsp tcsr; do
  SPMV_B200_LIB=$L/libspmv_b200_trace.so timeout 300 python tools/trace_rs.py $v c2 > $O/c25_trace_$v.log 2>&1; echo "trace $v rc=$?"
  SPMV_B200_LIB=$L/libspmv_b200_trace_early.so timeout 300 python tools/trace_rs.py $v c2 > $O/c25_trace_early_$v.log 2>&1; echo "trace early $v rc=$?"
done
SPMV_B200_LIB=$L/libspmv_b200_rsearly.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small or config or edge or options or batch or panels or csc or relu" > $O/c25_pytest.log 2>&1; echo "pytest(early) rc=$?"
tail -3 $O/c25_pytest.log
for rep in 1 2; do for lib in "" _rsearly; do for v in awsp tcsr; do for cfg in c2 c0 c1 c3; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py $v $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c25_rs.log
done; done; done; done

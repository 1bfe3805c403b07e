#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
for rep in 1 2; do for lib in "" _ast32 _ast24 _rmin48; do for cfg in c0 c3 c2; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py asp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c37_asp.log
done; done; done
SPMV_B200_LIB=$L/libspmv_b200_ast32.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "small or config or edge or options or batch or relu or asp_register" > $O/c37_pytest.log 2>&1; echo "pytest(ast32) rc=$?"
tail -3 $O/c37_pytest.log

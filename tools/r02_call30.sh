#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
SPMV_B200_LIB=$L/libspmv_b200_aspr32.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small or config or edge or options or relu" > $O/c30_pytest.log 2>&1; echo "pytest(aspr32) rc=$?"
tail -3 $O/c30_pytest.log
for rep in 1 2; do for lib in "" _aspr16 _aspr24 _aspr32; do for cfg in c2 c0 c1 c3; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py asp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c30_asp.log
done; done; done

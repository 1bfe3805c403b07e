#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "powerlaw or config4 or full_size or csc" > $O/c27_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c27_pytest.log
for lib in "" _m3c4 _m4c2 _m6c2 _m8c1 _m2c8; do for bc in 4 8 16; do
  SPMV_WSP_BIN_CTAS=$bc SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 300 python tools/c4_powerlaw.py 2>&1 | tail -1 | sed "s/^/[bin_ctas $bc] /" | tee -a $O/c27_c4.log
done; done

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small or config or edge or options or batch or panels or csc or relu or asp_register" > $O/c35_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c35_pytest.log
for cfg in c2 c0 c1 c3; do for geom in auto 3,8 3,7 2,7 2,8 3,6; do for lib in "" _st12; do
  if [ $geom = auto ]; then unset SPMV_WSP_RING_GEOM; else export SPMV_WSP_RING_GEOM=$geom; fi
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py wsp $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib geom $geom] /" | tee -a $O/c35_wsp.log
done; done; done

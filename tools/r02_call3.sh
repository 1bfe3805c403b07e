#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "strips or smoke" > $O/c3_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c3_pytest.log
SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200_ring.so timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c3_ring.log 2>&1; cat $O/c3_ring.log
timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c3_plain_strips.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strips_kernel -s 3 -c 1 -o $O/r02_strips_v3 python tools/c5_slab.py chunk_mode=4 > $O/c3_ncu_strips.log 2>&1
echo "ncu rc=$?"
cat $O/c3_plain_strips.log

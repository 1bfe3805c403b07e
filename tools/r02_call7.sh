#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c7_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/c7_pytest.log
for cfg in c2 c0 c1 c3; do
  case $cfg in c2) rs=0,8;; *) rs=0,32;; esac
  timeout 300 python tools/sweep.py asp $cfg row_splits=$rs >> $O/c7_asp.log 2>&1
done
cat $O/c7_asp.log
SPMV_SEED=1234 SPMV_STRICT=1 timeout 300 ./build/sparse_sgemv > $O/c7_harness.log 2>&1; echo "harness rc=$?"
tail -30 $O/c7_harness.log
timeout 300 python tools/sweep.py asp c2 row_splits=0 > $O/c7_plain_asp.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:asp_warp_kernel -s 3 -c 1 -o $O/r02_asp_v1 python tools/sweep.py asp c2 row_splits=0 > $O/c7_ncu.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c4_pytest.log
SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200_st32.so timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c4_st32.log 2>&1; cat $O/c4_st32.log
timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c4_st16.log 2>&1; cat $O/c4_st16.log
for cfg in c2 c0 c1 c3; do
  case $cfg in c2) rs=0,8;; *) rs=0,32;; esac
  timeout 300 python tools/sweep.py asp $cfg row_splits=$rs >> $O/c4_asp.log 2>&1
done
cat $O/c4_asp.log
SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200_st32.so timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c4_plain.log 2>&1 && \
SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200_st32.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:strips_kernel -s 3 -c 1 -o $O/r02_strips_v4 python tools/c5_slab.py chunk_mode=4 > $O/c4_ncu.log 2>&1
echo "ncu rc=$?"

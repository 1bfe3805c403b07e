#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c5_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c5_pytest.log
for v in "" _nopf _st16pf; do
  SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$v.so timeout 300 python tools/c5_slab.py chunk_mode=4 2>&1 | sed "s/^/[lib$v] /" | tee -a $O/c5_strips.log
done
for st in 16 32; do for ct in 2 3 4 5; do
  for cfg in c2 c0; do
    SPMV_ASP_STAGES=$st SPMV_ASP_CTAS=$ct timeout 200 python tools/sweep.py asp $cfg row_splits=0 2>&1 | sed "s/^/[st=$st ctas=$ct] /" | tee -a $O/c5_asp.log
  done
done; done
timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c5_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strips_kernel -s 3 -c 1 -o $O/r02_strips_v5 python tools/c5_slab.py chunk_mode=4 > $O/c5_ncu.log 2>&1
echo "ncu rc=$?"

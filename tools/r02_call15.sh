#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "config or small or options or edge or batch or relu" > $O/c15_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c15_pytest.log
for lib in "" _b8d5c3 _b8d6c3 _b8d8c2 _b4d5c3 _b4d6c3; do for v in awsp; do for cfg in c2 c0 c3; do
  SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 200 python tools/sweep.py $v $cfg slab_cols=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c15_panel.log
done; done; done
SPMV_PANEL_RS=0 timeout 200 python tools/sweep.py awsp c3 slab_cols=0 2>&1 | sed "s/^/[ring] /"

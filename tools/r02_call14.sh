#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c14_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/c14_pytest.log
for rs in 1 0 1 0; do for v in awsp tcsr; do for cfg in c2 c0 c1 c3; do
  SPMV_PANEL_RS=$rs timeout 200 python tools/sweep.py $v $cfg slab_cols=0 2>&1 | sed "s/^/[rs=$rs] /" | tee -a $O/c14_panel.log
done; done; done
timeout 300 python tools/sweep.py awsp c2 slab_cols=0 > $O/c14_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_rs_kernel -s 3 -c 1 -o $O/r02_panel_rs_v1 python tools/sweep.py awsp c2 slab_cols=0 > $O/c14_ncu.log 2>&1
echo "ncu rc=$?"

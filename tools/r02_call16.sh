#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c16_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/c16_pytest.log
for rs in 1 0; do for v in awsp tcsr; do for cfg in c2 c0 c3; do
  SPMV_PANEL_RS=$rs timeout 200 python tools/sweep.py $v $cfg slab_cols=0 2>&1 | sed "s/^/[rs=$rs] /" | tee -a $O/c16_panel.log
done; done; done

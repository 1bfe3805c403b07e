#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
SPMV_ASP_TMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "small or edge or asp_register or relu or column_slab or empty or options or config" > $O/c45_pytest.log 2>&1; echo "pytest(tma) rc=$?"
tail -3 $O/c45_pytest.log
for cfg in c2 c0 c3; do for t in 0 1; do
  SPMV_ASP_TMA=$t timeout 100 python tools/sweep.py asp $cfg index_bits=0 2>&1 | sed "s/^/[tma $t] /" | tee -a $O/c45_asp.log
done; done

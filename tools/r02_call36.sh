#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c36_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c36_pytest.log
for gpt in 8 4 2; do for cfg in c1 c0 c2; do
  SPMV_WSP_GROUPS_PER_THREAD=$gpt timeout 200 python tools/sweep.py wsp $cfg index_bits=0 2>&1 | sed "s/^/[gpt $gpt] /" | tee -a $O/c36_wsp.log
done; done
for gpt in 8 4; do SPMV_WSP_GROUPS_PER_THREAD=$gpt timeout 300 python tools/c4_powerlaw.py 2>&1 | tail -1 | sed "s/^/[gpt $gpt] /"; done
timeout 300 python tools/batch_bench.py c2 wsp 2>&1 | tail -4

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for cfg in c1 c3 c0; do
  timeout 200 python tools/sweep.py awsp $cfg chunk_mode=0,1,2,4 2>&1 | tee -a $O/c23_modes.log
  timeout 200 python tools/sweep.py tcsr $cfg chunk_mode=0,1,2 2>&1 | tee -a $O/c23_modes.log
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/c23_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/c23_pytest.log

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c12_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/c12_pytest.log
timeout 600 python tools/c4_powerlaw.py 2>&1 | tail -1
timeout 300 python tools/c5_slab.py chunk_mode=4 2>&1 | tail -1
for v in wsp asp awsp tcsr; do timeout 300 python tools/batch_bench.py c2 $v 2>&1 | tail -4; done

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q > $O/c46_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c46_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1

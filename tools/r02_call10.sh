#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
SPMV_PANEL_INTERLEAVE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "config or small or options or edge" > $O/c10_pytest.log 2>&1; echo "pytest(interleave) rc=$?"
tail -3 $O/c10_pytest.log
for il in 0 1 0 1; do for v in awsp tcsr; do for cfg in c2 c0 c3; do
  SPMV_PANEL_INTERLEAVE=$il timeout 200 python tools/sweep.py $v $cfg slab_cols=0 2>&1 | sed "s/^/[il=$il] /" | tee -a $O/c10_panel.log
done; done; done
timeout 900 python -m pytest tests -m gpu -x -q -k "device_packer" > $O/c10_pytest2.log 2>&1; echo "pytest(device packer) rc=$?"; tail -3 $O/c10_pytest2.log

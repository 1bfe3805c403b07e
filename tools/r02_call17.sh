#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for rep in 1 2; do for lib in "" _nobb _c4d4 ring; do for cfg in c2 c0 c3; do
  if [ "$lib" = ring ]; then
    SPMV_PANEL_RS=0 timeout 200 python tools/sweep.py awsp $cfg slab_cols=0 2>&1 | sed "s/^/[ring] /" | tee -a $O/c17_panel.log
  else
    SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 200 python tools/sweep.py awsp $cfg slab_cols=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c17_panel.log
  fi
done; done; done

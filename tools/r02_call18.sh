#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for rep in 1 2; do for lib in _nobb _nobbpad _nobbpadcv _bbcv; do for cfg in c2 c0; do
    SPMV_B200_LIB=$PWD/spmv_test_b200/lib/libspmv_b200$lib.so timeout 200 python tools/sweep.py awsp $cfg slab_cols=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c18_panel.log
done; done; done

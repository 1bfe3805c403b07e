#!/bin/bash
# Round 2, GPU call 1: full GPU suite with the row-strip kernel, the launch / streaming floor,
# the config-5 slab in its three forms, one ncu capture of the strips kernel.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c1_pytest.log
(cd tools/ubench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/floor.bin floor.cu) > $O/c1_floor.log 2>&1 && timeout 300 /tmp/floor.bin >> $O/c1_floor.log 2>&1; echo "floor rc=$?"
cat $O/c1_floor.log
timeout 600 python tools/c5_slab.py chunk_mode=4,3,0 > $O/c1_c5.log 2>&1; echo "c5 rc=$?"
cat $O/c1_c5.log
timeout 300 python tools/c5_slab.py chunk_mode=4 > $O/c1_plain_strips.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strips_kernel -s 3 -c 2 -o $O/r02_strips_v1 python tools/c5_slab.py chunk_mode=4 > $O/c1_ncu_strips.log 2>&1
echo "ncu rc=$?"
tail -3 $O/c1_ncu_strips.log

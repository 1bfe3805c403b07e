#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
(cd tools/ubench && nvcc -O3 -std=c++17 --extended-lambda -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4.bin gather4.cu) > $O/c44_build.log 2>&1; echo "build rc=$?"
for br in 1 4; do
  timeout 90 /tmp/gather4.bin $br 8 > $O/c44_gather4_box$br.log 2>&1; echo "box_rows $br rc=$?"; cat $O/c44_gather4_box$br.log
done
good=1; grep -q "^gather4 box" $O/c44_gather4_box1.log || good=4
for stg in 4 12; do timeout 90 /tmp/gather4.bin $good $stg 2>&1 | tail -1 | tee -a $O/c44_gather4_stages.log; done
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
cap() {  # tag, kernel regex, command...
  local tag=$1 rx=$2; shift 2
  timeout 400 "$@" > $O/c39_plain_$tag.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/r02f_$tag "$@" > $O/c39_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"; tail -1 $O/c39_plain_$tag.log
}
cap wsp_c2 wsp_ring_kernel python tools/sweep.py wsp c2 index_bits=0
cap wsp_c1 wsp_ring_kernel python tools/sweep.py wsp c1 index_bits=0
cap wsp_c4 wsp_merged python tools/c4_powerlaw.py
timeout 900 python bench.py --quick --steps 4 --warmup 3 > $O/c39_plain_quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_step_final.csv python bench.py --quick --steps 4 --warmup 3 > $O/c39_ncu_launches.log 2>&1
echo "launch list rc=$?"

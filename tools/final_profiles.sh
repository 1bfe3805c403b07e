#!/bin/bash
# Round-end evidence run (one GPU): plain runs first, ncu only after the same command exited 0.
set -u
mkdir -p gpurun_out
O=gpurun_out
run_ncu() {  # tag, kernel regex, command...
  local tag=$1 rx=$2; shift 2
  "$@" > $O/plain_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 2 -o $O/r01f_$tag "$@" > $O/ncu_$tag.log 2>&1
  echo "$tag: rc=$?"
}
run_ncu wsp wsp_ring python bench.py --quick --variant wsp --steps 20 --warmup 3
run_ncu asp asp_kernel python bench.py --quick --variant asp --steps 20 --warmup 3
run_ncu awsp panel_kernel python bench.py --quick --variant awsp --steps 20 --warmup 3
run_ncu tcsr panel_kernel python bench.py --quick --variant tcsr --steps 20 --warmup 3
run_ncu c5lob panel_kernel python tools/c5_slab.py chunk_mode=3
run_ncu c5row panel_kernel python tools/c5_slab.py chunk_mode=0
python bench.py --no-aux --steps 20 --warmup 3 > $O/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01f_launches_step.csv \
    python bench.py --no-aux --steps 20 --warmup 3 > $O/ncu_step.log 2>&1
echo "launch list rc=$?"
ls -la $O | tail -20

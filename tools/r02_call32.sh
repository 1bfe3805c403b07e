#!/bin/bash
# Round-2 evidence run (one GPU): tests, smoke, the full bench, the reference arm, then ncu (each only after the same
# command exited 0 without it).
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c32_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c32_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c32_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/c32_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/c32_bench.json 2> $O/c32_bench.err; echo "bench rc=$?"; tail -3 $O/c32_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/c32_ref.json 2> $O/c32_ref.err; echo "ref rc=$?"; cut -c1-400 $O/c32_ref.json
timeout 900 python bench.py --quick --steps 4 --warmup 3 > $O/c32_plain_quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_step_final.csv python bench.py --quick --steps 4 --warmup 3 > $O/c32_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {  # tag, kernel regex, command...
  local tag=$1 rx=$2; shift 2
  timeout 400 "$@" > $O/c32_plain_$tag.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/r02f_$tag "$@" > $O/c32_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
}
cap asp_c2 asp_kernel python tools/sweep.py asp c2 index_bits=0
cap awsp_c2 panel_rs_kernel python tools/sweep.py awsp c2 index_bits=0
cap tcsr_c2 panel_rs_kernel python tools/sweep.py tcsr c2 index_bits=0
cap wsp_c4 wsp_merged python tools/c4_powerlaw.py
cap strips_c5 strips_kernel python tools/c5_slab.py chunk_mode=4
ls -la $O/*.ncu-rep | tail

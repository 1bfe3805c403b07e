#!/bin/bash
# evidence run: launch list of the bench command, config-4 wsp capture, final-ish full bench
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --quick --steps 4 --warmup 3 > $O/c11_plain_quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_step.csv python bench.py --quick --steps 4 --warmup 3 > $O/c11_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 python tools/c4_powerlaw.py > $O/c11_plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wsp_merged -s 3 -c 1 -o $O/r02_c4_wsp python tools/c4_powerlaw.py > $O/c11_ncu_c4.log 2>&1
echo "c4 ncu rc=$?"; tail -3 $O/c11_plain_c4.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/c11_bench.json 2> $O/c11_bench.err; echo "bench rc=$?"; tail -3 $O/c11_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/c11_ref.json 2> $O/c11_ref.err; echo "ref rc=$?"; cat $O/c11_ref.json | cut -c1-600

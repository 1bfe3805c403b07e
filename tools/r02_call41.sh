#!/bin/bash
# Round-2 evidence run, final N = 1 pass (after asp's per-plan kernel choice): tests, smoke, full bench, reference arm, asp capture.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c41_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c41_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c41_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/c41_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/c41_bench.json 2> $O/c41_bench.err; echo "bench rc=$?"; tail -3 $O/c41_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/c41_ref.json 2> $O/c41_ref.err; echo "ref rc=$?"; cut -c1-300 $O/c41_ref.json
timeout 400 python tools/sweep.py asp c2 index_bits=0 > $O/c41_plain_asp.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:asp_kernel -s 3 -c 1 -o $O/r02f_asp_c2 python tools/sweep.py asp c2 index_bits=0 > $O/c41_ncu_asp.log 2>&1
echo "ncu asp rc=$?"; tail -1 $O/c41_plain_asp.log
for cfg in c0 c1 c3; do timeout 200 python tools/sweep.py asp $cfg index_bits=0 2>&1 | tail -1; done
timeout 300 python tools/c4_powerlaw.py 2>&1 | tail -1; timeout 200 python tools/sweep.py wsp c1 index_bits=0 2>&1 | tail -1

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c13_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/c13_pytest.log
timeout 300 python tools/c5_slab.py chunk_mode=4 2>&1 | tail -1
timeout 600 python tools/c4_powerlaw.py 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --quick > $O/c13_bench_quick.json 2> $O/c13_bench.err; echo "bench rc=$?"; tail -3 $O/c13_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c13_bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','parity_ok') if k in d}, d['roofline']['us_per_launch'], d['roofline']['frac'])
PY

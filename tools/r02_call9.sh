#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c9_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/c9_pytest.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 --quick > $O/c9_bench_quick.json 2> $O/c9_bench.err; echo "bench rc=$?"; tail -3 $O/c9_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9_bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','roofline','build','parity_ok') if k in d})
PY

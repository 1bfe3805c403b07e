#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > $O/c43_pytest_mg.log 2>&1; echo "pytest mg rc=$?"
tail -4 $O/c43_pytest_mg.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c43_bench_n2.json 2> $O/c43_bench_n2.err; echo "bench n2 rc=$?"
tail -2 $O/c43_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c43_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','n_gpus','join','parity_ok') if k in d}, d['e2e'].get('us_per_step'))
print({k:v for k,v in d['parity'].items() if k!='oracle'})
PY

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -k "small or config or edge or options or batch or panels or csc or relu or group" > $O/c26_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c26_pytest.log
for rep in 1 2; do for lib in "" _rse1 _rse2 _rse3; do for v in awsp tcsr; do for cfg in c2 c0 c1 c3; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 200 python tools/sweep.py $v $cfg index_bits=0 2>&1 | sed "s/^/[lib$lib] /" | tee -a $O/c26_rs.log
done; done; done; done
timeout 900 python bench.py --steps 20 --warmup 5 --quick > $O/c26_bench_quick.json 2> $O/c26_bench.err; echo "bench rc=$?"; tail -3 $O/c26_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c26_bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','parity_ok') if k in d}, d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e'])
PY

#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -k "strips or full_size or group or device_packer" > $O/c40_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c40_pytest.log
for rep in 1 2 3; do for lib in "" _ma1; do
  SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 300 python tools/c5_slab.py chunk_mode=4 2>&1 | tail -1 | sed "s/^/[lib$lib] /"
done; done
timeout 900 python bench.py --steps 20 --warmup 5 --quick > $O/c40_bench_quick.json 2> $O/c40_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c40_bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','parity_ok') if k in d}, d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['us_per_step'])
PY

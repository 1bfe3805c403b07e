#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
L=$PWD/spmv_test_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q > $O/c28_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c28_pytest.log
for lib in "" _m3c2 _m5c2 _m4c3; do for bc in 2 3 4 5 6; do
  SPMV_WSP_BIN_CTAS=$bc SPMV_B200_LIB=$L/libspmv_b200$lib.so timeout 300 python tools/c4_powerlaw.py 2>&1 | tail -1 | sed "s/^/[bin_ctas $bc] /" | tee -a $O/c28_c4.log
done; done
timeout 300 python tools/c5_slab.py chunk_mode=4 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --quick > $O/c28_bench_quick.json 2> $O/c28_bench.err; echo "bench rc=$?"; tail -3 $O/c28_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c28_bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','us_per_step','parity_ok') if k in d}, d['roofline']['us_per_launch'], d['roofline']['frac'], d['e2e']['us_per_step'])
PY

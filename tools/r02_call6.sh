#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
(cd tools/ubench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rowpattern.bin rowpattern.cu) > $O/c6_rowpattern.log 2>&1 && timeout 300 /tmp/rowpattern.bin >> $O/c6_rowpattern.log 2>&1; echo "rowpattern rc=$?"
cat $O/c6_rowpattern.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc=$?"
tail -5 $O/c6_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c6_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','us_per_step','roofline','e2e','join','parity','build','parity_ok','us_per_step_stats') if k in d})
    for c,v in d.get('configs',{}).items():
        print(c, {k:(vv.get('us_per_call'),vv.get('hot_l2_us_per_call'),vv.get('frac_alg'),vv.get('frac_phys'),vv.get('parity'),vv.get('l2_copies')) for k,vv in v.items() if isinstance(vv,dict)})
    print(d.get('gpu_comparator')); print(d.get('cpu_baseline')); print(d.get('cpu_baseline_c2'))
except Exception as e: print('parse failed', e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > $O/c6_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c6_pytest.log

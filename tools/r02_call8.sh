#!/bin/bash
# two GPUs: the group tests (IPC, symmetric memory, in-process), then the bench at N = 2 with both joins
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/c8_smi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > $O/c8_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/c8_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c8_bench_n2.json 2> $O/c8_bench_n2.err; echo "bench n2 rc=$?"
tail -5 $O/c8_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --join nccl > $O/c8_bench_n2_nccl.json 2> $O/c8_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
tail -3 $O/c8_bench_n2_nccl.err
SPMV_SEED=1234 SPMV_STRICT=1 timeout 300 ./build/sparse_sgemv 2>&1 | tail -4
python - <<'PY'
import json
for f in ('c8_bench_n2.json','c8_bench_n2_nccl.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','us_per_step','n_gpus','join','parity_ok','e2e') if k in d})
        print(d['parity'])
    except Exception as e: print(f,'parse failed',e)
PY

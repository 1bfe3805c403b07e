#!/bin/bash
# eight GPUs: the bench at N = 8 and N = 4 (arrival join), N = 8 with the NCCL join, the harness over 8 devices
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > $O/c34_bench_n$n.json 2> $O/c34_bench_n$n.err; echo "bench n$n rc=$?"
  tail -2 $O/c34_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 5 --join nccl > $O/c34_bench_n8_nccl.json 2> $O/c34_bench_n8_nccl.err; echo "bench n8 nccl rc=$?"
SPMV_SEED=1234 SPMV_STRICT=1 timeout 300 ./build/sparse_sgemv 2>&1 | tail -3
python - <<'PY'
import json
for f in ('c34_bench_n8.json','c34_bench_n4.json','c34_bench_n8_nccl.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','us_per_step','n_gpus','join','parity_ok') if k in d}, d['e2e'].get('us_per_step'), d['roofline']['us_per_launch'])
        print('   ', {k:v for k,v in d['parity'].items() if k!='oracle'})
    except Exception as e: print(f,'parse failed',e)
PY

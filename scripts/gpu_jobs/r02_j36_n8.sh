#!/bin/bash
# Round 2, job 36 (8 GPUs): the bench line at N = 8 with the final build (value, parity_check, e2e; no lean / dense legs).
mkdir -p gpurun_out
O=gpurun_out/r02_j36
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-lean-leg --fill 0 ) > $O.bench_n8.json 2> $O.bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_j36.bench_n8.json').read().strip().splitlines()[-1])
    print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),'frac',d['roofline']['frac'], d['e2e']['where_the_time_goes'])
except Exception as e:
    print('failed',e); print(open('gpurun_out/r02_j36.bench_n8.err').read()[-3000:])
PY

#!/bin/bash
# Round 2, job 39 (2 GPUs): parity_check with vacuum row-strips on real ranks.
mkdir -p gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-lean-leg --fill 0 ) > gpurun_out/r02_j39.bench_n2.json 2> gpurun_out/r02_j39.bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_j39.bench_n2.json').read().strip().splitlines()[-1])
    print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),d['parity_detail'])
except Exception as e:
    print('failed',e); print(open('gpurun_out/r02_j39.bench_n2.err').read()[-3000:])
PY

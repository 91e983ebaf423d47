#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j12
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > $O.bench_n2.json 2> $O.bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_j12.bench_n2.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),json.dumps(d.get('parity_detail')),'lean',d['lean_interior']['value'],'dense',d['dense']['value'],d['dense']['ratio_to_value'])
except Exception as e:
    print('failed',e); print(open('gpurun_out/r02_j12.bench_n2.err').read()[-3000:])
PY

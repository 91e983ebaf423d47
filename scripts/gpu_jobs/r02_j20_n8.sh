#!/bin/bash
# Round 2, job 20 (8 GPUs): bench line at N = 8 and N = 1 on the same box (e2e as redefined: palette eps in, far-field table out).
mkdir -p gpurun_out
O=gpurun_out/r02_j20
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 20 --warmup 5 ) > $O.bench_n8.json 2> $O.bench_n8.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg ) > $O.bench_n1.json 2> $O.bench_n1.err
for n in 8 1; do python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_j20.bench_n%s.json'%n).read().strip().splitlines()[-1])
    print('N',n,'value',d['value'],'e2e',d['e2e']['value'],d['e2e']['with_field_snapshot']['value'],'parity',d.get('parity_check'),'lean',d['lean_interior']['value'],'dense',d['dense']['value'], d['e2e']['where_the_time_goes'])
except Exception as e:
    print('N',n,'failed',e); print(open('gpurun_out/r02_j20.bench_n%s.err'%n).read()[-3000:])
PY
done

#!/bin/bash
# Round 2, job 17 (8 GPUs): the benchmark line with its parity_check at N = 8 and N = 4.
mkdir -p gpurun_out
O=gpurun_out/r02_j17
nvidia-smi -L > $O.gpus.txt; nvidia-smi topo -m > $O.topo.txt 2>&1
for n in 8 4; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $n --steps 20 --warmup 5 ) > $O.bench_n$n.json 2> $O.bench_n$n.err
done
for n in 8 4; do python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_j17.bench_n%s.json'%n).read().strip().splitlines()[-1])
    print('N',n,'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),d.get('parity_detail',{}).get('ntff_uw_rel_err_after_reduce'),'lean',d['lean_interior']['value'],'dense',d['dense']['value'], d['e2e']['host_placement'])
except Exception as e:
    print('N',n,'failed',e); print(open('gpurun_out/r02_j17.bench_n%s.err'%n).read()[-3000:])
PY
done

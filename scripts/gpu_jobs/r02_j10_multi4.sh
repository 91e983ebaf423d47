#!/bin/bash
# Round 2, job 10 (4 GPUs): multi-rank bit-identity (middle ranks have both neighbours) + bench at N = 2, 4.
mkdir -p gpurun_out
O=gpurun_out/r02_j10
nvidia-smi -L > $O.gpus.txt
( timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -k "4-TM_UPML_2D-peer-f64-fused or 4-TE_UPML_2D-peer-f64-fused or 4-TM_UPML_2D-peer-f64-leanfused or 4-TE_UPML_2D-peer-f64-exact or 4-TM_UPML_2D-nccl-f64-exact or 4-TE_UPML_2D-peer-f64-leanfused or 2-TM_UPML_2D-peer-f64-fused or 4-TM_UPML_2D-peer-f64-unit" 2>&1 | tail -15 ) > $O.pytest_multi.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_peer_local.py tests/test_gpu_plugin_devices.py -x -q 2>&1 | tail -8 ) > $O.pytest_local_spread.log 2>&1
for n in 4 2; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $n --steps 20 --warmup 5 ) > $O.bench_n$n.json 2> $O.bench_n$n.err
done
tail -n 12 $O.pytest_multi.log $O.pytest_local_spread.log
for n in 4 2; do python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_j10.bench_n%s.json'%n).read().strip().splitlines()[-1])
    print('N',n,'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),d.get('parity_detail',{}).get('ntff_uw_rel_err_after_reduce'),'lean',d['lean_interior']['value'],'dense',d['dense']['value'], d['e2e']['host_placement'])
except Exception as e:
    print('N',n,'failed',e); print(open('gpurun_out/r02_j10.bench_n%s.err'%n).read()[-3000:])
PY
done

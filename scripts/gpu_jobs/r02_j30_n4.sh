#!/bin/bash
# Round 2, job 30 (4 GPUs), final build: 4-rank bit-identity cases (two ranks with both neighbours), the C plugin's slabs
# spread over four devices (ids 2-5), the bench line at N = 4, BASELINE configs[2] at N = 4.
mkdir -p gpurun_out
O=gpurun_out/r02_j30
nvidia-smi -L > $O.gpus.log
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "4-TM_UPML_2D-peer-f64-fused or 4-TE_UPML_2D-peer-f64-fused or 4-TM_UPML_2D-peer-f64-leanfused or 4-TE_UPML_2D-peer-f64-exact or 4-TM_UPML_2D-nccl-f64-exact" ) > $O.pytest_multi.log 2>&1
tail -n 5 $O.pytest_multi.log
( time timeout 600 python -m pytest tests/test_gpu_plugin_devices.py -x -q -m gpu ) > $O.pytest_plugin.log 2>&1
tail -n 5 $O.pytest_plugin.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 4 --steps 20 --warmup 5 ) > $O.bench_n4.json 2> $O.bench_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_j30.bench_n4.json').read().strip().splitlines()[-1])
    print('N 4 value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),'lean',d['lean_interior']['value'],'dense',d['dense']['value'])
except Exception as e:
    print('N 4 failed',e); print(open('gpurun_out/r02_j30.bench_n4.err').read()[-3000:])
PY
bash scripts/gpu_jobs/r02_j29_config2.sh 4

#!/bin/bash
# Round 2, job 28 (2 GPUs): the multi-GPU tests with the final build (torchrun 2-rank cases, the C plugin's slabs on two
# devices, ids 4/5 included), the bench line at N = 2, BASELINE configs[2] (8192^2, strong scaling) at N = 2.
mkdir -p gpurun_out
O=gpurun_out/r02_j28
nvidia-smi -L > $O.gpus.log
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_plugin_devices.py tests/test_gpu_peer_local.py -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 6 $O.pytest.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > $O.bench_n2.json 2> $O.bench_n2.err
for model in LAYER MORPHO_SCALE; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 2 --steps 20 --warmup 5 --model $model --n 8192 --strong --no-lean-leg ) > $O.config2_${model}_n2.json 2> $O.config2_${model}_n2.err
done
python - <<'PY'
import json
for f in ('bench_n2','config2_LAYER_n2','config2_MORPHO_SCALE_n2'):
    try:
        d=json.loads(open('gpurun_out/r02_j28.%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),'scaling',d.get('scaling'),d['config']['workload'][:70])
    except Exception as e:
        print(f,'failed',e); print(open('gpurun_out/r02_j28.%s.err'%f).read()[-2000:])
PY

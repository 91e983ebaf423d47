#!/bin/bash
# Round 2, job 34 (G GPUs): final build on several GPUs: peer-halo tests (G = 2: the torchrun 2-rank cases, G = 4: five
# 4-rank cases), the C plugin's slabs, the bench line at N = G (parity_check included), N = 1 on the same box.
G=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_j34_n$G
if [ "$G" = "2" ]; then
  ( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_plugin_devices.py -x -q -m gpu ) > $O.pytest.log 2>&1
else
  ( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "4-TM_UPML_2D-peer-f64-fused or 4-TE_UPML_2D-peer-f64-fused or 4-TM_UPML_2D-peer-f64-leanfused or 4-TE_UPML_2D-peer-f64-exact or 4-TM_UPML_2D-nccl-f64-exact" ) > $O.pytest.log 2>&1
fi
tail -n 5 $O.pytest.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $G --steps 20 --warmup 5 ) > $O.bench.json 2> $O.bench.err
if [ "$G" = "2" ]; then ( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg --no-ntff-leg ) > $O.bench_n1.json 2> $O.bench_n1.err; fi
python - $G <<'PY'
import json, sys
G = sys.argv[1]
for f in ('bench', 'bench_n1'):
    try:
        d=json.loads(open('gpurun_out/r02_j34_n%s.%s.json'%(G,f)).read().strip().splitlines()[-1])
        print(f,'N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),'lean',d['lean_interior']['value'],'dense',d['dense']['value'],'frac',d['roofline']['frac'], d['e2e']['where_the_time_goes']['h2d_s'])
    except Exception as e:
        print(f,'failed',e); print(open('gpurun_out/r02_j34_n%s.%s.err'%(G,f)).read()[-2000:])
PY

#!/bin/bash
# Round 2, job 29 (G GPUs, G = $1): BASELINE configs[2] -- 8192 x 8192, multiLayerModel / morphoScaleModel, strong scaling.
# (--size, not --n: torchrun's argparse rejects "--n" as an ambiguous abbreviation of its own options.)
G=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02_j29
for model in LAYER MORPHO_SCALE; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $G --steps 20 --warmup 5 --model $model --size 8192 --strong --no-lean-leg ) > $O.config2_${model}_n$G.json 2> $O.config2_${model}_n$G.err
done
python - $G <<'PY'
import json, sys
G = sys.argv[1]
for f in ('config2_LAYER_n' + G, 'config2_MORPHO_SCALE_n' + G):
    try:
        d=json.loads(open('gpurun_out/r02_j29.%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity_check'),'scaling',d.get('scaling'),d['config']['workload'][:90])
    except Exception as e:
        print(f,'failed',e); print(open('gpurun_out/r02_j29.%s.err'%f).read()[-2000:])
PY

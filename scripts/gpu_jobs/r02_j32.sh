#!/bin/bash
# Round 2, job 32: vacuum row-strips of the one-pass step (no E arrays kept where eps == 1 along a tile row).
mkdir -p gpurun_out
O=gpurun_out/r02_j32
( time timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "vacuum or replay_from or default_on_large" ) > $O.pytest_new.log 2>&1
tail -n 25 $O.pytest_new.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg --no-ntff-leg ) > $O.bench.json 2> $O.bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_j32.bench.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('value',d['value'],'e2e',d['e2e']['value'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'])
    print('roofline frac',r['frac'],'ms',r['ms_per_launch'],'bytes/cell',r['algorithmic_bytes_per_cell'],r['algorithmic_bytes_per_cell_by_region'])
    print('lean kernel',d['lean_interior'].get('kernel'), d['lean_interior']['moved_bytes_per_cell_update'])
except Exception as e:
    print('bench failed',e); print(open('gpurun_out/r02_j32.bench.err').read()[-3000:])
PY

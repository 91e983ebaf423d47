#!/bin/bash
# Round 2, job 14: NS TE block-uniform interior form, 4096^2 direct oracle comparison, dense bench with the reciprocal cache.
mkdir -p gpurun_out
O=gpurun_out/r02_j14
( timeout 900 python -m pytest tests/test_gpu_split.py -x -q 2>&1 | tail -8 ) > $O.pytest_split.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "4096" 2>&1 | tail -8 ) > $O.pytest_4096.log 2>&1
( timeout 600 python scripts/split_bench.py 4096 7 ) > $O.split_bench7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split_te -s 40 -c 2 -o $O.split_te -f \
    python scripts/split_bench.py 4096 7 > $O.ncu_split.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $O.bench.json 2> $O.bench.err
tail -n 8 $O.pytest_split.log $O.pytest_4096.log; cat $O.split_bench7.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j14.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'],d['dense']['ratio_to_value'],'e2e',d['e2e']['value'],'plugin',d['e2e_plugin']['value'],d['e2e_plugin']['init_s'])
PY

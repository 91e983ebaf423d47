#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j11
( timeout 900 python -m pytest tests/test_gpu_peer_local.py -x -q -k random 2>&1 | tail -25 ) > $O.pytest_random.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8 ) > $O.pytest.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg ) > $O.bench.json 2> $O.bench.err
tail -n 25 $O.pytest_random.log; tail -n 8 $O.pytest.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j11.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'],d['dense']['ratio_to_value'],'e2e',d['e2e']['value'])
PY

#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j19
( timeout 900 python -m pytest tests/test_gpu_dropin_main.py tests/test_gpu_split.py tests/test_gpu_replay.py -x -q 2>&1 | tail -6 ) > $O.pytest.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $O.bench.json 2> $O.bench.err
tail -n 6 $O.pytest.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j19.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['with_field_snapshot'])
print('plugin',json.dumps(d['e2e_plugin']))
PY

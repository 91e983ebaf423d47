#!/bin/bash
# Round 2, job 9: exact reciprocal-based permittivity division + pulse shortcut: parity + dense throughput.
mkdir -p gpurun_out
O=gpurun_out/r02_j9
( timeout 1800 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_lean.py tests/test_gpu_f32.py tests/test_gpu_plugin_devices.py -x -q 2>&1 | tail -15 ) > $O.pytest.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $O.bench.json 2> $O.bench.err
tail -n 15 $O.pytest.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j9.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'],d['dense']['ratio_to_value'],'e2e',d['e2e']['value'],'plugin',d['e2e_plugin'])
PY

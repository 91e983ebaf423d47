#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j18
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ntff_entry.py tests/test_gpu_sweep.py tests/test_gpu_dropin_main.py tests/test_gpu_plugin_devices.py -x -q 2>&1 | tail -12 ) > $O.pytest.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 ) > $O.bench.json 2> $O.bench.err
tail -n 12 $O.pytest.log; tail -n 3 $O.bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j18.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'],'traffic',d['roofline']['traffic'])
print('e2e',json.dumps(d['e2e']))
print('plugin',json.dumps(d['e2e_plugin']))
PY

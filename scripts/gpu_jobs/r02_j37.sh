#!/bin/bash
# Round 2, job 37: the driver's default bench invocation once more on a fresh box (job 33's host copied at 1.4 GB/s during the e2e leg).
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_j37.bench_n1_tm.json 2> gpurun_out/r02_j37.bench_n1_tm.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j37.bench_n1_tm.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['where_the_time_goes']['h2d_s'],'plugin',d['e2e_plugin']['value'],d['e2e_plugin']['finish_s'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'],'lean',d['lean_interior']['value'],'dense',d['dense']['value'])
PY

#!/bin/bash
# Round 2, job 13: NS TE interior form (tests, timing, ncu of the split kernels); dense bench after the reciprocal cache;
# TE bench line.
mkdir -p gpurun_out
O=gpurun_out/r02_j13
( timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_dropin_main.py -x -q 2>&1 | tail -8 ) > $O.pytest_split.log 2>&1
( timeout 600 python scripts/split_bench.py 4096 ) > $O.split_bench.log 2>&1
( MPIFDTD_SPLIT_DENSE=1 timeout 600 python scripts/split_bench.py 4096 7 ) > $O.split_bench_dense7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:split_te -s 40 -c 10 -o $O.split_te -f \
    python scripts/split_bench.py 4096 7 > $O.ncu_split.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-plugin-leg --solver TE_UPML_2D ) > $O.bench_te.json 2> $O.bench_te.err
tail -n 8 $O.pytest_split.log; cat $O.split_bench.log $O.split_bench_dense7.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j13.bench_te.json').read().strip().splitlines()[0])
print('TE value',d['value'],'lean',d['lean_interior']['value'],'e2e',d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'])
PY

#!/bin/bash
# Round 2, job 22: occupancy A/B of the split kernels; bench line with the NTFF projection leg.
mkdir -p gpurun_out
O=gpurun_out/r02_j22
cp mpifdtd_b200/libmpifdtd_b200.so /tmp/lib_default.so
for lib in /tmp/lib_default.so build/variants/lib_split_mb5.so build/variants/lib_split_mb6.so build/variants/lib_split_mb8.so; do
  cp $lib mpifdtd_b200/libmpifdtd_b200.so
  echo "== $lib" >> $O.split.log
  ( timeout 600 python scripts/split_bench.py 4096 2>&1 | grep solver_id | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['solver_id'], '%.4f ms  %.2f Gcell/s' % (d['ms_per_step'], d['gcell_updates_per_s']))" ) >> $O.split.log 2>&1
done
cp /tmp/lib_default.so mpifdtd_b200/libmpifdtd_b200.so
cat $O.split.log
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $O.bench.json 2> $O.bench.err
tail -n 3 $O.bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_j22.bench.json').read().strip().splitlines()[0])
print('value',d['value'],'e2e',d['e2e']['value'],'ntff',json.dumps(d.get('ntff_projection')))
PY

#!/bin/bash
# Round 2, job 25: deferred stepping of the split-field solvers; their throughput at 4096^2 and on small grids.
mkdir -p gpurun_out
O=gpurun_out/r02_j25
( time timeout 900 python -m pytest tests/test_gpu_replay.py tests/test_gpu_sweep.py tests/test_gpu_split.py -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 15 $O.pytest.log
for n in 256 1024 4096; do
  for defer in 0 1; do
    echo "== n=$n MPIFDTD_DEFER_STEPS=$defer" >> $O.split.log
    ( MPIFDTD_DEFER_STEPS=$defer timeout 600 python scripts/split_bench.py $n 2>&1 | grep solver_id | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['solver_id'], '%.4f ms  %.2f Gcell/s' % (d['ms_per_step'], d['gcell_updates_per_s']))" ) >> $O.split.log 2>&1
  done
done
cat $O.split.log

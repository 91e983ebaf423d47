#!/bin/bash
# Round 2, job 26: deferred stepping of the split-field solvers, steady state (graph built during warm-up).
mkdir -p gpurun_out
O=gpurun_out/r02_j26
( time timeout 900 python -m pytest tests/test_gpu_replay.py tests/test_gpu_sweep.py tests/test_gpu_split.py -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 6 $O.pytest.log
for n in 256 512 1024; do
  for defer in 0 1; do
    echo "== n=$n MPIFDTD_DEFER_STEPS=$defer (warm-up 256, 1024 timed steps)" >> $O.split.log
    ( SPLIT_BENCH_STEPS=1024 SPLIT_BENCH_WARMUP=256 MPIFDTD_DEFER_STEPS=$defer timeout 600 python scripts/split_bench.py $n 2>&1 | grep solver_id | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['solver_id'], '%.4f ms  %.2f Gcell/s' % (d['ms_per_step'], d['gcell_updates_per_s']))" ) >> $O.split.log 2>&1
  done
done
cat $O.split.log

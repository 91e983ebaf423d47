#!/bin/bash
# Round 2, job 1: stress of the one-pass TMA kernel + sanitizer passes.
mkdir -p gpurun_out
O=gpurun_out/r02_j1
nvidia-smi -L > $O.gpu.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -5 ) > $O.pytest_fused.log 2>&1
# the failing test's own concurrency: register / cp.async / TMA forms side by side
( timeout 900 python scripts/fused_stress.py small 300 700 48 0,0,10,13,20,23,22 3000 5 ) > $O.small_testlike.log 2>&1
# bulk: default shape 20 and shape 23, > 1e5 launches each
( timeout 1200 python scripts/fused_stress.py small 300 700 48 20,23 17000 25 ) > $O.small_bulk.log 2>&1
for c in "45 47 7" "64 1030 1" "257 66 256" "131 200 64" "300 41 33"; do
  ( timeout 600 python scripts/fused_stress.py small $c 20,23 2000 25 ) >> $O.small_ragged.log 2>&1
done
( timeout 1500 python scripts/fused_stress.py large 16384 16384 20 10000 0 ) > $O.large_16384.log 2>&1
( timeout 600 python scripts/fused_stress.py large 16001 16411 20 1500 1 ) > $O.large_ragged_storeh.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck python scripts/fused_stress.py small 300 700 48 20,23 3 1 ) > $O.memcheck.log 2>&1
( timeout 900 compute-sanitizer --tool synccheck python scripts/fused_stress.py small 300 700 48 20,23 3 1 ) > $O.synccheck.log 2>&1
tail -n 3 $O.*.log

#!/bin/bash
# Round 2, job 24: angle batches of the split-field solvers; compute-sanitizer over the one-pass kernels.
mkdir -p gpurun_out
O=gpurun_out/r02_j24
( time timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_split.py -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 15 $O.pytest.log
rm -f $O.sanitizer.log
for tool in memcheck racecheck synccheck; do
  for solver in TM_UPML_2D TE_UPML_2D; do
    for lean in 0 1; do
      echo "== $tool $solver lean=$lean" >> $O.sanitizer.log
      ( timeout 420 compute-sanitizer --tool $tool --print-limit 5 python scripts/onepass_one.py $solver $lean 20 32 1024 3 MIE_CYLINDER 2>&1 | grep -v "^$" | tail -n 8 ) >> $O.sanitizer.log 2>&1
    done
  done
done
cat $O.sanitizer.log

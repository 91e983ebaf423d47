#!/bin/bash
# Round 2, job 6: ncu --set full of the one-pass kernels (TM lean, TM exact, TE exact) at 16384^2.
mkdir -p gpurun_out
O=gpurun_out/r02_j6
for cfg in "TM_UPML_2D 1 tm_lean" "TM_UPML_2D 0 tm_exact" "TE_UPML_2D 0 te_exact"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -s 2 -c 1 \
      -o $O.$3 -f python scripts/onepass_one.py $1 $2 20 32 16384 4 > $O.$3.log 2>&1
done
ls -la gpurun_out/r02_j6*

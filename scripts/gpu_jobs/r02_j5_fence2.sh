#!/bin/bash
# Round 2, job 5: slot release with a register dependency (mode 2) vs the proxy fence (mode 0).
mkdir -p gpurun_out
O=gpurun_out/r02_j5
for mode in 2 0; do
  touch mpifdtd_b200/csrc/engine/fused_kernels.cu
  make -s -C mpifdtd_b200/csrc PTXAS_V=-DB200_RELEASE_MODE=$mode > $O.build$mode.log 2>&1
  ( timeout 900 python scripts/fused_repro.py 200 2 ) > $O.mode$mode.repro_tm.log 2>&1
  ( timeout 600 python scripts/onepass_bench.py 16384 ZIGZAG TM_UPML_2D,TE_UPML_2D quick ) > $O.mode$mode.bench.log 2>&1
  tail -n 2 $O.mode$mode.repro_tm.log; cat $O.mode$mode.bench.log
done

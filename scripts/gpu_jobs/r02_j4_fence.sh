#!/bin/bash
# Round 2, job 4: the slot-release fix, two realisations A/B: correctness (fresh-engine repro loop) and cost.
mkdir -p gpurun_out
O=gpurun_out/r02_j4
for mode in 0 1; do
  if [ $mode = 1 ]; then
    touch mpifdtd_b200/csrc/engine/fused_kernels.cu
    make -s -C mpifdtd_b200/csrc PTXAS_V=-DB200_RELEASE_MODE=1 > $O.build1.log 2>&1
  fi
  ( timeout 900 python scripts/fused_repro.py 150 2 ) > $O.mode$mode.repro_tm.log 2>&1
  ( timeout 600 python scripts/fused_repro.py 60 3 ) > $O.mode$mode.repro_te.log 2>&1
  ( timeout 600 python scripts/onepass_bench.py 16384 ZIGZAG TM_UPML_2D,TE_UPML_2D quick ) > $O.mode$mode.bench.log 2>&1
  tail -n 3 $O.mode$mode.repro_tm.log $O.mode$mode.repro_te.log; cat $O.mode$mode.bench.log
done

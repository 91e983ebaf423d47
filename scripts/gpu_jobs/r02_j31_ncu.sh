#!/bin/bash
# Round 2, job 31: ncu --set full of the split-field kernels (ids 0, 1, 6, 7 at 4096^2, final launch bounds) and of the
# NTFF kernels (sample, projection, spectrum) of BASELINE configs[1] (1024^2, 2000 steps).
mkdir -p gpurun_out
O=gpurun_out/r02_j31
for id in 0 1 6 7; do
  MPIFDTD_DEFER_STEPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:split_t -s 40 -c 2 \
      -o $O.split_id$id -f python scripts/split_bench.py 4096 $id > $O.split_id$id.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntff_project -c 1 -o $O.ntff_project -f \
    python scripts/config2_full.py --no-ref > $O.ntff_project.log 2>&1
MPIFDTD_DEFER_STEPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntff_sample -s 1500 -c 1 -o $O.ntff_sample -f \
    python scripts/config2_full.py --no-ref > $O.ntff_sample.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum -c 3 -o $O.ntff_spectrum -f \
    python scripts/config2_full.py --no-ref > $O.ntff_spectrum.log 2>&1
ls -la gpurun_out/r02_j31*

#!/bin/bash
# Launch-bounds sweep of the TE UPML kernels on the GPU box (it has nvcc): rebuild
# upml_kernels.cu with -DB200_TE_E_MIN_BLOCKS=<n> and time the TE step at 16384^2.
set -e
cd "$(dirname "$0")/.."
for e in 1 4 5 6; do
  touch mpifdtd_b200/csrc/engine/upml_kernels.cu
  make -s -C mpifdtd_b200/csrc PTXAS_V="-DB200_TE_E_MIN_BLOCKS=$e" > /dev/null 2>&1
  echo "TE_E_MIN_BLOCKS=$e $(python bench.py --solver TE_UPML_2D --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['roofline']['e_phase']['ms_per_launch'], d['roofline']['ms_per_launch'])")"
done

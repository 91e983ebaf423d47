#!/bin/bash
# A/B of the pipelined step against the two-kernel step at the bench size, over band heights
# and occupancy targets (rebuilds upml_kernels.cu on the box).
cd "$(dirname "$0")/.."
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(round(d['value'],2), round(d['ms_per_step'],3))"; }
echo "two-kernel TM: $(B200FDTD_PIPELINED=0 run)"
for mb in 3 4; do
  touch mpifdtd_b200/csrc/engine/upml_kernels.cu
  make -s -C mpifdtd_b200/csrc PTXAS_V="-DB200_PIPE_MIN_BLOCKS=$mb" > /dev/null 2>&1
  for rows in 2 4 8; do
    echo "pipelined TM min_blocks=$mb band_rows=$rows: $(B200FDTD_PIPELINED=1 B200FDTD_PIPE_BAND_ROWS=$rows run)"
  done
done
echo "pipelined TE min_blocks=4 band_rows=4: $(B200FDTD_PIPELINED=1 B200FDTD_PIPE_BAND_ROWS=4 run --solver TE_UPML_2D)"

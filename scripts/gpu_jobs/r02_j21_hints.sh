#!/bin/bash
# Round 2, job 21: cache-hint A/B of the one-pass kernels (builds made on the CPU box: build/variants/lib_s<store>l<load>.so)
mkdir -p gpurun_out
O=gpurun_out/r02_j21
cp mpifdtd_b200/libmpifdtd_b200.so /tmp/lib_default.so
for lib in /tmp/lib_default.so build/variants/lib_s1l0.so build/variants/lib_s0l1.so build/variants/lib_s1l1.so /tmp/lib_default.so; do
  cp $lib mpifdtd_b200/libmpifdtd_b200.so
  echo "== $lib" >> $O.log
  ( timeout 600 python scripts/onepass_bench.py 16384 ZIGZAG TM_UPML_2D,TE_UPML_2D quick 2>&1 | grep "shape 20 band  32" ) >> $O.log 2>&1
done
cp /tmp/lib_default.so mpifdtd_b200/libmpifdtd_b200.so
cat $O.log

#!/bin/bash
# Round 2, job 35: launch shapes / band heights of the one-pass kernels with the vacuum row-strips.
mkdir -p gpurun_out
( timeout 500 python scripts/onepass_bench.py 16384 ZIGZAG ) > gpurun_out/r02_j35.shapes.log 2>&1
cat gpurun_out/r02_j35.shapes.log | grep -v "^$" | tail -n 50

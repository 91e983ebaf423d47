#!/bin/bash
# Round 2, job 42: compact ring for tiles inside the lean rectangle (lean one-pass kernels only; the exact kernels' SASS is unchanged).
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_lean.py -x -q -m gpu -k "lean" ) > gpurun_out/r02_j42.pytest.log 2>&1
tail -n 6 gpurun_out/r02_j42.pytest.log
( timeout 120 python scripts/onepass_bench.py 16384 ZIGZAG TM_UPML_2D,TE_UPML_2D quick ) > gpurun_out/r02_j42.shapes.log 2>&1
grep "lean\|exact shape 20 band  32" gpurun_out/r02_j42.shapes.log

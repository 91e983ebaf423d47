#!/bin/bash
# Round 2, job 2: the rewritten one-pass kernel family (TM/TE x exact/lean): tests, stress, timing.
mkdir -p gpurun_out
O=gpurun_out/r02_j2
( timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15 ) > $O.pytest_fused.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_peer_local.py -x -q 2>&1 | tail -30 ) > $O.pytest_peer_local.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_lean.py tests/test_gpu_unit.py tests/test_gpu_replay.py -x -q 2>&1 | tail -15 ) > $O.pytest_lean_unit.log 2>&1
( timeout 900 python scripts/onepass_bench.py 16384 ZIGZAG ) > $O.onepass_bench.log 2>&1
( timeout 600 python scripts/fused_stress.py small 300 700 48 20,23,21,24 3000 25 ) > $O.small_bulk.log 2>&1
tail -n 40 $O.pytest_fused.log $O.pytest_peer_local.log $O.pytest_lean_unit.log $O.onepass_bench.log $O.small_bulk.log

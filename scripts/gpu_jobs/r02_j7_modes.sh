#!/bin/bash
# Round 2, job 7: specialised consumer loops (ROW_UNIT / ROW_LEAN / ROW_GENERAL): tests + timing.
mkdir -p gpurun_out
O=gpurun_out/r02_j7
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_peer_local.py -x -q 2>&1 | tail -15 ) > $O.pytest.log 2>&1
( timeout 600 python scripts/fused_repro.py 100 2 ) > $O.repro_tm.log 2>&1
( timeout 600 python scripts/onepass_bench.py 16384 ZIGZAG TM_UPML_2D,TE_UPML_2D quick ) > $O.bench.log 2>&1
tail -n 15 $O.pytest.log; tail -n 2 $O.repro_tm.log; cat $O.bench.log

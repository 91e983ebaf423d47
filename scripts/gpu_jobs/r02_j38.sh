#!/bin/bash
# Round 2, job 38: peer halos through vacuum row-strips (slabs of whole 256-column strips), the edge-kernel refresh fix.
mkdir -p gpurun_out
O=gpurun_out/r02_j38
( time timeout 900 python -m pytest tests/test_gpu_peer_local.py -x -q -m gpu -k "vacuum_row_strips" ) > $O.pytest.log 2>&1
tail -n 12 $O.pytest.log

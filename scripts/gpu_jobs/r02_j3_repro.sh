#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j3
( timeout 1200 python scripts/fused_repro.py 150 2 ) > $O.repro_tm.log 2>&1
( timeout 900 python scripts/fused_repro.py 60 3 ) > $O.repro_te.log 2>&1
tail -n 60 $O.repro_tm.log $O.repro_te.log

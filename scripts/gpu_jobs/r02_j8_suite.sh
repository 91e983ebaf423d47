#!/bin/bash
# Round 2, job 8: whole GPU suite + default bench line at N = 1.
mkdir -p gpurun_out
O=gpurun_out/r02_j8
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O.pytest.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O.bench.json 2> $O.bench.err
tail -n 30 $O.pytest.log; cat $O.bench.json; tail -n 5 $O.bench.err

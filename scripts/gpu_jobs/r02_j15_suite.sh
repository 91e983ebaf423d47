#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_j15
( time timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > $O.pytest.log 2>&1
( time timeout 600 python -c "import __graft_entry__ as g; print(g.smoke())" ) > $O.smoke.log 2>&1
tail -n 30 $O.pytest.log; tail -n 6 $O.smoke.log

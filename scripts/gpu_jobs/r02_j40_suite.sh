#!/bin/bash
# Round 2, job 40: the whole GPU suite + smoke with the final library.
mkdir -p gpurun_out
O=gpurun_out/r02_j40
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $O.pytest.log 2>&1
tail -n 6 $O.pytest.log
( time timeout 200 python -c "import __graft_entry__ as g; print(g.smoke())" ) > $O.smoke.log 2>&1
grep "smoke\[" $O.smoke.log | cut -c1-160

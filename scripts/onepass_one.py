"""A few launches of ONE configuration of the one-pass step, for ncu:
    python scripts/onepass_one.py SOLVER LEAN SHAPE BAND [N] [REPS] [MODEL]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun

solver, lean, shape, band = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 16384
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
model = sys.argv[7] if len(sys.argv) > 7 else "ZIGZAG"
run = SlabRun(model, solver, n, n, 64, with_ntff=False)
e = run.engine
run.L.mpifdtd_upml_step_args(run.kind, 0, B.C.byref(run.args))
e.set_option(B.OPT_FUSED, 1)
e.set_option(B.OPT_LEAN_INTERIOR, lean)
e.set_option(B.OPT_FUSED_SHAPE, shape)
e.set_option(B.OPT_BAND_ROWS, band)
for _ in range(reps):
    e.phase_fused(run.args)
e.sync()
run.close()

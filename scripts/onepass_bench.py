"""Kernel-level timing of the one-pass step (pre-pass + marching kernel, b200fdtd_phase_fused) at
benchmark scale: TM / TE x exact / lean x launch shapes x band heights.  CUDA events on the engine's
stream.  `python scripts/onepass_bench.py [n] [model]` prints one line per configuration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = sys.argv[2] if len(sys.argv) > 2 else "ZIGZAG"
solvers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["TM_UPML_2D", "TE_UPML_2D"]
quick = len(sys.argv) > 4 and sys.argv[4] == "quick"
BYTES = {("TM_UPML_2D", 0): 232, ("TM_UPML_2D", 1): 136, ("TE_UPML_2D", 0): 272, ("TE_UPML_2D", 1): 176}
reps = 10
for solver in solvers:
    run = SlabRun(model, solver, n, n, 64, with_ntff=False)
    e = run.engine
    run.L.mpifdtd_upml_step_args(run.kind, 0, B.C.byref(run.args))
    for lean in (0, 1):
        e.set_option(B.OPT_LEAN_INTERIOR, lean)
        # (bands above 63 rows do not fit the 64-bit row masks of the vacuum row-strips: E arrays kept everywhere)
        for shape, band in (((20, 32), (21, 32), (20, 63)) if quick else
                            ((20, 32), (21, 32), (24, 32), (23, 32), (22, 32), (20, 48), (21, 48), (20, 63), (21, 63),
                             (20, 16), (20, 64))):
            e.set_option(B.OPT_FUSED_SHAPE, shape)
            e.set_option(B.OPT_BAND_ROWS, band)
            try:
                e.phase_fused(run.args); e.phase_fused(run.args); e.sync()
            except B.EngineError as err:
                print(solver, "lean" if lean else "exact", "shape", shape, "band", band, "unavailable:", str(err)[:80])
                continue
            e.timer_start()
            for _ in range(reps):
                e.phase_fused(run.args)
            ms = e.timer_stop() / reps
            by = BYTES[(solver, lean)] - (40 if run.kind == 2 else 80) * e.vacuum_cells() / float(n * n)
            print("%s %-5s shape %d band %3d: %7.3f ms  %6.2f Gcell/s  %6.0f GB/s (%.1f B/cell)"
                  % (solver, "lean" if lean else "exact", shape, band, ms, n * n / ms / 1e6, by * n * n / ms / 1e6, by),
                  flush=True)
    run.close()

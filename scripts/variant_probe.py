"""Run a few steps of each TM step form (for ncu per-kernel metrics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200.slab import SlabRun
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for cfg in [dict(B200FDTD_FUSED="0", B200FDTD_STORE_H="1"), dict(B200FDTD_FUSED="0", B200FDTD_STORE_H="0"),
            dict(B200FDTD_FUSED="1", B200FDTD_STORE_H="0", B200FDTD_BAND_ROWS="64")]:
    os.environ.update(cfg)
    run = SlabRun("ZIGZAG", "TM_UPML_2D", n, n, 4, with_ntff=False)
    for _ in range(3):
        run.step()
    run.engine.sync()
    run.close()

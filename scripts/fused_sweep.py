"""Tuning sweep of the fused TM step (launch shape, band height, H stores) on one GPU."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200.slab import SlabRun

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = 10
configs = [dict(B200FDTD_FUSED="0", B200FDTD_STORE_H="1"), dict(B200FDTD_FUSED="0", B200FDTD_STORE_H="0")]
shapes = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [20, 21, 24, 22, 25, 23]
bands = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [256, 1024]
for shape, band, store in itertools.product(shapes, bands, [0]):
    configs.append(dict(B200FDTD_FUSED="1", B200FDTD_FUSED_SHAPE=str(shape), B200FDTD_BAND_ROWS=str(band),
                        B200FDTD_STORE_H=str(store)))

for cfg in configs:
    for k in ("B200FDTD_FUSED", "B200FDTD_FUSED_SHAPE", "B200FDTD_BAND_ROWS", "B200FDTD_STORE_H"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    run = SlabRun("ZIGZAG", "TM_UPML_2D", n, n, K + 3, with_ntff=False)
    for _ in range(3):
        run.step()
    run.engine.sync()
    run.engine.timer_start()
    for _ in range(K):
        run.step()
    ms = run.engine.timer_stop() / K
    print("%-90s %.3f ms/step  %.2f Gcell/s" % (cfg, ms, n * n / ms / 1e6), flush=True)
    run.close()

import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
from mpifdtd_b200 import binding as B
L = B.lib()
os.chdir("/tmp")
for defer in ("0", "1"):
    os.environ["MPIFDTD_DEFER_STEPS"] = defer
    for batch in (None, list(range(0, 185, 5))):
        gpu = B.Plugin("MIE_CYLINDER", 2, 256, steps=2000, angle_batch=batch)
        h = gpu.engine_handle()
        L.b200fdtd_sync(h)
        t0 = time.perf_counter()
        gpu.run()
        t1 = time.perf_counter()
        h = gpu.engine_handle(); L.b200fdtd_sync(h)
        t2 = time.perf_counter()
        print("defer", defer, "batch", len(batch) if batch else 1, "issue %.3f s  drained %.3f s  launches %d" % (t1 - t0, t2 - t0, gpu.launches()), flush=True)
        gpu.finish()

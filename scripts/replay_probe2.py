import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MPIFDTD_TIMING"] = "1"
from scripts.sweep_bench import sweep
sweep(128, 50, 0, 10, 5, 0)
for defer in ("0", "1"):
    os.environ["MPIFDTD_DEFER_STEPS"] = defer
    sys.stderr.write("==== defer %s batched 256\n" % defer)
    t = time.perf_counter(); sweep(256, 2000, 0, 180, 5, 0); sys.stderr.write("total %.3f s\n" % (time.perf_counter() - t))

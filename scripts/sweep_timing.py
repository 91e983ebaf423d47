import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["MPIFDTD_TIMING"] = "1"
from scripts.sweep_bench import sweep
sweep(128, 50, 0, 10, 5, 0)
for n, steps in ((256, 2000), (1024, 1000)):
    sys.stderr.write("==== %d batched\n" % n)
    t = time.perf_counter(); sweep(n, steps, 0, 180, 5, 0); sys.stderr.write("total %.3f s\n" % (time.perf_counter() - t))
    sys.stderr.write("==== %d one at a time\n" % n)
    t = time.perf_counter(); sweep(n, steps, 0, 180, 5, 1); sys.stderr.write("total %.3f s\n" % (time.perf_counter() - t))

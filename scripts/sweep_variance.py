"""Repeat the batched 1024^2 sweep a few times with the init/finish timers on (stderr)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MPIFDTD_TIMING"] = "1"
from scripts.sweep_bench import sweep
sweep(128, 50, 0, 10, 5, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for rep in range(4):
    sys.stderr.write("==== rep %d\n" % rep)
    t = time.perf_counter(); sweep(n, 1000, 0, 180, 5, 0); sys.stderr.write("total %.3f s\n" % (time.perf_counter() - t))

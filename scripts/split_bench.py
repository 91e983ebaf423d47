#!/usr/bin/env python
"""Throughput of the split-field solvers (ids 0, 1: Yee + Berenger PML; 6, 7: NS-FDTD) at
BASELINE configs[3]'s size, 4096 x 4096, Mie cylinder, through the plugin surface
(simulator_calc).  CUDA-event timed on the engine's stream; one JSON line per solver id.
Algorithmic bytes per cell-update: lean form (default for ids 0, 1, 6):
id 0 216 B, id 1 224 B, id 6 224 B; dense form (id 7, or MPIFDTD_SPLIT_DENSE=1): 5 complex
fields read/written + 8 coefficients + source factors = 264-288 B."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200 import binding as B

DENSE = os.environ.get("MPIFDTD_SPLIT_DENSE") == "1"
# id 0: H phase reads Ezx,Ezy,Hx,Hy (64) + writes Hx,Hy (32); E phase reads Hx,Hy,Ezx,Ezy (64) + eps (8) + writes 3 (48)
# id 1: E phase reads Hzx,Hzy,Ex,Ey (64) + 2 eps (16) + writes 2 (32); H phase reads 4 (64) + writes 3 (48)
# id 6: H phase reads Ez,Hx,Hy (48) + 2 numerators (16) + writes 2 (32); E phase reads 4 (64) + numerator + source
#       factor (16) + writes 3 (48)
# id 7 (dense): H phase reads 4 + 4 coef (96) + writes 3 (48); E phase reads Hz,Ex,Ey (48) + 4 coef + 2 factors (48) + writes 2 (32)
# id 7, interior form (default): H phase reads 4 + 1 coef (72) + writes 3 (48); E phase reads Hz,Ex,Ey (48) + 2 coef + 2
#       factors (32) + writes 2 (32) = 232; the 10-cell frame keeps the dense 272
BYTES = {0: 280, 1: 288, 6: 264, 7: 272} if DENSE else {0: 216, 1: 224, 6: 224, 7: 232}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    K, W = int(os.environ.get("SPLIT_BENCH_STEPS", 40)), int(os.environ.get("SPLIT_BENCH_WARMUP", 5))
    L = B.lib()
    kinds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 6, 7]
    for kind in kinds:
        gpu = B.Plugin("MIE_CYLINDER", kind, n, steps=K + W + 2, lambda_nm=500)
        h = gpu.engine_handle()
        gpu.step(W)
        h = gpu.engine_handle()              # taking the handle hands the recorded steps to the engine
        B.check(L.b200fdtd_sync(h), "sync")
        B.check(L.b200fdtd_timer_start(h), "timer_start")
        gpu.step(K)
        gpu.engine_handle()
        ms = C.c_float(0)
        B.check(L.b200fdtd_timer_stop(h, C.byref(ms)), "timer_stop")
        rate = n * n * K / (ms.value * 1e-3) / 1e9
        print(json.dumps({"solver_id": kind, "n": n, "steps": K, "ms_per_step": ms.value / K,
                          "gcell_updates_per_s": rate, "algorithmic_bytes_per_cell_update": BYTES[kind],
                          "achieved_GBs": rate * BYTES[kind], "frac_of_measured_6451": rate * BYTES[kind] / 6451.2}))
        sys.stdout.flush()
        cwd = os.getcwd()
        import tempfile
        os.chdir(tempfile.mkdtemp())
        L.mpifdtd_setAngleBatch(None, 0)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        gpu.finish()
        os.dup2(saved, 1)
        os.chdir(cwd)


if __name__ == "__main__":
    main()

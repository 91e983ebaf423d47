#!/usr/bin/env python
"""Throughput of the split-field solvers (ids 0, 1: Yee + Berenger PML; 6, 7: NS-FDTD) at
BASELINE configs[3]'s size, 4096 x 4096, Mie cylinder, through the plugin surface
(simulator_calc).  CUDA-event timed on the engine's stream; one JSON line per solver id.
Algorithmic bytes per cell-update: 5 complex fields, 8 dense coefficients and one
or two source factors: TM 280 B, TE 288 B."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpifdtd_b200 import binding as B

BYTES = {0: 280, 6: 280 + 0, 1: 288, 7: 288}     # NS TM reads Ez (1 field) instead of Ezx+Ezy in the H phase: 264
BYTES[6] = 264
BYTES[7] = 272


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    K, W = 40, 5
    L = B.lib()
    for kind in (0, 1, 6, 7):
        gpu = B.Plugin("MIE_CYLINDER", kind, n, steps=K + W + 2, lambda_nm=500)
        h = gpu.engine_handle()
        gpu.step(W)
        B.check(L.b200fdtd_sync(h), "sync")
        B.check(L.b200fdtd_timer_start(h), "timer_start")
        gpu.step(K)
        ms = C.c_float(0)
        B.check(L.b200fdtd_timer_stop(h, C.byref(ms)), "timer_stop")
        rate = n * n * K / (ms.value * 1e-3) / 1e9
        print(json.dumps({"solver_id": kind, "n": n, "steps": K, "ms_per_step": ms.value / K,
                          "gcell_updates_per_s": rate, "algorithmic_bytes_per_cell_update": BYTES[kind],
                          "achieved_GBs": rate * BYTES[kind], "frac_of_measured_6546": rate * BYTES[kind] / 6546.2}))
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

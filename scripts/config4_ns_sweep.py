#!/usr/bin/env python
"""BASELINE configs[3]: NS-FDTD (nsFdtdTM / nsFdtdTE), concentric-circle model, 4096 x 4096,
wavelength sweep; far field per wavelength = the one-shot frequency-domain NTFF
(ntffTM_Frequency, ntffTM.c:72-158) of the final fields, as SURVEY 8(d) prescribes (the NS
coefficients depend on k, so one run per wavelength).  The concentric model ships disabled
upstream (models.c:82-90); MPIFDTD_ENABLE_CONCENTRIC=1 opts in.

    python scripts/config4_ns_sweep.py [n] [steps] [lambda_first] [lambda_last] [lambda_step]
prints one JSON line per wavelength and a summary line."""
import ctypes as C
import json
import os
import sys
import tempfile
import time

os.environ["MPIFDTD_ENABLE_CONCENTRIC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mpifdtd_b200 import binding as B


def main():
    a = [int(x) for x in sys.argv[1:]]
    n, steps, l0, l1, dl = (a + [4096, 1000, 400, 700, 50][len(a):])[:5]
    L = B.lib()
    os.chdir(tempfile.mkdtemp(prefix="config4_"))
    devnull = os.open(os.devnull, os.O_WRONLY)
    total_t, total_updates = 0.0, 0
    for solver, kind in (("NS_TM_2D", 6), ("NS_TE_2D", 7)):
        for lam in range(l0, l1 + 1, dl):
            saved = os.dup(1); os.dup2(devnull, 1)
            t0 = time.perf_counter()
            gpu = B.Plugin("CONCENTRIC_CIRCLE", kind, n, steps=steps, lambda_nm=lam)
            t_init = time.perf_counter() - t0
            h = gpu.engine_handle()
            B.check(L.b200fdtd_timer_start(h), "timer_start")
            gpu.run()
            gpu.engine_handle()              # hands the recorded steps to the engine
            ms = C.c_float(0)
            B.check(L.b200fdtd_timer_stop(h, C.byref(ms)), "timer_stop")
            far = np.zeros(360, dtype=np.complex128)
            if kind == 6:
                L.mpifdtd_ntffFrequency(kind, far.ctypes.data)
            gpu.finish()
            wall = time.perf_counter() - t0
            os.dup2(saved, 1); os.close(saved)
            total_t += wall; total_updates += n * n * steps
            print(json.dumps({"solver": solver, "lambda_nm": lam, "n": n, "steps": steps,
                              "init_s": t_init, "stepping_ms": ms.value, "wall_s": wall,
                              "gcell_updates_per_s_stepping": n * n * steps / (ms.value * 1e-3) / 1e9,
                              "far_field_max": float(np.abs(far).max()),
                              "far_field_l2": float(np.sqrt((np.abs(far) ** 2).sum()))}))
            sys.stdout.flush()
    print(json.dumps({"summary": "BASELINE configs[3] sweep", "runs": 2 * len(range(l0, l1 + 1, dl)),
                      "wall_s": total_t, "gcell_updates_per_s_wall": total_updates / total_t / 1e9}))


if __name__ == "__main__":
    main()

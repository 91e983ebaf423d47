"""Reproduction loop for the gating test's large case with FRESH engines every repetition (the
stress loop re-uses its engines and never saw a mismatch; the test creates them anew).  After every
step the digests of all nine arrays are compared; the first differing (step, engine, array) is
located cell by cell.

    python scripts/fused_repro.py [REPS] [KIND] [NPX NPY BAND]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from mpifdtd_b200 import binding as B
from test_gpu_fused import make_engine, random_case

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
npx, npy, band = (int(x) for x in sys.argv[3:6]) if len(sys.argv) > 5 else (300, 700, 48)
NAMES = {2: ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"], 3: ["Ex", "Jx", "Dx", "Ey", "Jy", "Dy", "Hz", "Mz", "Bz"]}[kind]
steps = 6
L = B.lib()
L.models_setModel(B.MODELS["NO_MODEL"])
L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
eps, state = random_case(npx, npy, kind)
SPEC = [(0, 1, 20), (0, 0, 20), (1, 0, 20), (1, 1, 20), (1, 1, 23), (1, 0, 23), (1, 1, 24), (1, 0, 21), (1, 0, 22)]
bad_total = 0
for rep in range(reps):
    engines = [make_engine(L, npx, npy, steps, eps, fused, store_h=sh, band=band, shape=shape, kind=kind)
               for fused, sh, shape in SPEC]
    for eng in engines:
        for slot in range(9):
            eng.set_field(slot, state[slot])
    args = B.StepArgs()
    L.field_reset()
    found = False
    for step in range(steps):
        L.mpifdtd_upml_step_args(kind, 1, C.byref(args))
        for eng in engines:
            eng.step(args)
        L.field_nextStep()
        want = [engines[0].digest(s) for s in range(9)]
        for n, eng in enumerate(engines[1:], 1):
            for slot in range(9):
                if eng.digest(slot) != want[slot]:
                    a, b = engines[0].get_field(slot), eng.get_field(slot)
                    ii, jj = np.nonzero(a != b)
                    print("rep %d step %d engine %d %s array %s: %d cells; rows %s cols %s" %
                          (rep, step, n, SPEC[n], NAMES[slot], len(ii), sorted(set(ii.tolist()))[:40],
                           sorted(set(jj.tolist()))[:40]), flush=True)
                    for i, j in list(zip(ii, jj))[:12]:
                        print("    (%d,%d) row-in-band %d: want %r got %r" % (i, j, (i - 1) % band, a[i, j], b[i, j]))
                    found = True
                    break
        if found:
            break
    bad_total += found
    for eng in engines:
        eng.close()
print("FUSED_REPRO kind %d %dx%d band %d: %d repetitions with fresh engines, %d with a mismatch" %
      (kind, npx, npy, band, reps, bad_total), flush=True)

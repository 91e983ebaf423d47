"""Stress of the one-pass step: many repetitions of a small ragged case per launch shape, every
array compared bit for bit with the two-kernel step; prints where the first difference sits."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from mpifdtd_b200 import binding as B
from test_gpu_fused import make_engine

npx, npy, band = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (300, 700, 48)))
shapes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [20, 23, 22, 24]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 30
steps = 6
L = B.lib()
L.models_setModel(B.MODELS["NO_MODEL"])
L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
rng = np.random.default_rng(npx * 1000 + npy)
eps = np.where(rng.random((npx, npy)) < 0.5, 1.0, 1.0 + 2.0 * rng.random((npx, npy)))
state = [rng.standard_normal((npx, npy)) + 1j * rng.standard_normal((npx, npy)) for _ in range(9)]
mu0 = B.MU_0_S
state[3] = (state[5].real / mu0) + 1j * (state[5].imag / mu0)
state[6] = (state[8].real / mu0) + 1j * (state[8].imag / mu0)
engines = [make_engine(L, npx, npy, steps, eps, 0, store_h=0)]
engines += [make_engine(L, npx, npy, steps, eps, 1, store_h=(n % 2), band=band, shape=s) for n, s in enumerate(shapes)]
bad = {s: 0 for s in shapes}
for rep in range(reps):
    for eng in engines:
        for slot in range(9):
            eng.set_field(slot, state[slot])
    args = B.StepArgs()
    L.field_reset()
    for _ in range(steps):
        L.mpifdtd_upml_step_args(2, 1, C.byref(args))
        for eng in engines:
            eng.step(args)
        L.field_nextStep()
    ref = [engines[0].get_field(s) for s in range(9)]
    for s, eng in zip(shapes, engines[1:]):
        for slot in range(9):
            got = eng.get_field(slot)
            diff = got.view(np.float64).reshape(npx, npy, 2) != ref[slot].view(np.float64).reshape(npx, npy, 2)
            if diff.any():
                bad[s] += 1
                ii, jj = np.nonzero(diff.any(axis=2))
                print("rep", rep, "shape", s, "slot", slot, "cells", len(ii), "rows", ii.min(), ii.max(), "cols", jj.min(),
                      jj.max(), "first", (ii[0], jj[0]), flush=True)
                break
print("mismatching repetitions per shape:", bad)

"""Stress of the one-pass (TMA-staged) TM step against the two-kernel step, bit for bit.

Two modes, both compare on the DEVICE with b200fdtd_field_digest (no plane ever leaves the GPU
unless a mismatch has to be located):

  small  NPX NPY BAND SHAPES REPS [RESET_EVERY]   the gating test's case (tests/test_gpu_fused.py) with the
         test's concurrency: one two-kernel engine + one one-pass engine per launch shape, each on
         its own stream, all stepped side by side from the same random state; after EVERY step the
         digests of all nine arrays are compared.  A mismatch is located (step, array, rows,
         columns, band, row in band) and the whole repetition is re-run 100 x from the same
         start state to get a hit rate.
  large  NPX NPY SHAPE STEPS [STORE_H]  one two-kernel engine + one one-pass engine at benchmark
         scale from a random state and a random eps map; Ez digest compared after every launch,
         all nine arrays every 50.

Prints one summary line per mode: `FUSED_STRESS <mode> ... launches=<n> mismatches=<m>`.
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from mpifdtd_b200 import binding as B
from test_gpu_fused import make_engine

NAMES = ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"]


def locate(ref_eng, eng, slot, band, first_row=1):
    a, b = ref_eng.get_field(slot), eng.get_field(slot)
    diff = a.view(np.float64).reshape(a.shape + (2,)) != b.view(np.float64).reshape(a.shape + (2,))
    ii, jj = np.nonzero(diff.any(axis=2))
    if len(ii) == 0:
        return "digest differs but the planes compare equal on the host (transient?)"
    rows_in_band = sorted(set(int((i - first_row) % band) for i in ii))[:8]
    return ("%d cells, rows %d..%d (band %d..%d, row-in-band %s), columns %d..%d, first (%d, %d)"
            % (len(ii), ii.min(), ii.max(), (ii.min() - first_row) // band, (ii.max() - first_row) // band,
               rows_in_band, jj.min(), jj.max(), ii[0], jj[0]))


def small(npx, npy, band, shapes, reps, reset_every=25):
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
    engines += [make_engine(L, npx, npy, steps, eps, 1, store_h=(n % 2), band=band, shape=s)
                for n, s in enumerate(shapes)]
    args = B.StepArgs()

    def one_rep(reset=True, stop_at_first=True):
        """6 steps -> list of (step, shape, slot) mismatches; reset: from the random start state,
        else from wherever the previous repetition left the (identical) engines"""
        if reset:
            for eng in engines:
                for slot in range(9):
                    eng.set_field(slot, state[slot])
        L.field_reset()
        bad = []
        for step in range(steps):
            L.mpifdtd_upml_step_args(2, 1, C.byref(args))
            for eng in engines:                        # all launched before anybody is waited for
                eng.step(args)
            L.field_nextStep()
            want = [engines[0].digest(s) for s in range(9)]
            for shape, eng in zip(shapes, engines[1:]):
                for slot in range(9):
                    if eng.digest(slot) != want[slot]:
                        bad.append((step, shape, slot, eng))
                        break
            if bad and stop_at_first:
                break
        return bad

    launches, mism = 0, {s: 0 for s in shapes}
    t0 = time.time()
    for rep in range(reps):
        bad = one_rep(reset=(rep % reset_every == 0))
        launches += steps * len(shapes)
        for step, shape, slot, eng in bad:
            mism[shape] += 1
            print("MISMATCH rep %d step %d shape %d array %s: %s" % (rep, step, shape, NAMES[slot],
                                                                      locate(engines[0], eng, slot, band)), flush=True)
        if bad:
            hits = sum(1 for _ in range(100) if one_rep())
            print("  100 fresh repetitions from the random start state: %d mismatch" % hits, flush=True)
        if rep % 500 == 499:
            print("  rep %d, %.0f s" % (rep + 1, time.time() - t0), flush=True)
    print("FUSED_STRESS small %dx%d band %d shapes %s: reps=%d one-pass launches=%d (each compared with the "
          "two-kernel step on all 9 arrays) mismatches=%s" % (npx, npy, band, shapes, reps, launches, mism), flush=True)
    for eng in engines:
        eng.close()
    return sum(mism.values())


def large(npx, npy, shape, steps, store_h):
    L = B.lib()
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, min(steps, 8000)))
    rng = np.random.default_rng(7)
    eps = np.where(rng.random((npx, npy), dtype=np.float32) < 0.98, 1.0, 2.5)      # 2 % material cells
    z = rng.standard_normal((npx, npy), dtype=np.float32).astype(np.complex128)
    z += 1j * np.roll(z.real, 1, axis=1)
    mu0 = B.MU_0_S
    h = (z.real / mu0) + 1j * (z.imag / mu0)
    engines = []
    for fused in (0, 1):
        eng = B.Engine(2, npx, npy, 10)
        ti, tj = np.empty((6, npx)), np.empty((6, npy))
        L.mpifdtd_upml_tables(2, ti.ctypes.data, tj.ctypes.data)
        eng.set_tables(ti, tj)
        eng.set_eps(0, eps)
        eng.set_option(B.OPT_FUSED, fused)
        eng.set_option(B.OPT_STORE_H, store_h)
        if fused:
            eng.set_option(B.OPT_FUSED_SHAPE, shape)
        for slot in range(9):
            eng.set_field(slot, h if slot in (3, 6) else z)
        engines.append(eng)
    del z, h, eps
    ref, one = engines
    assert one.step_form() == 3 and ref.step_form() != 3
    args = B.StepArgs()
    L.field_reset()
    mism, t0 = 0, time.time()
    for step in range(steps):
        L.mpifdtd_upml_step_args(2, 0, C.byref(args))
        ref.step(args)
        one.step(args)
        L.field_nextStep()
        slots = range(9) if step % 50 == 49 or step == steps - 1 else (0,)
        for slot in slots:
            if ref.digest(slot) != one.digest(slot):
                mism += 1
                print("MISMATCH step %d array %s: %s" % (step, NAMES[slot], locate(ref, one, slot, 32)), flush=True)
                break
        if mism > 3:
            break
        if step % 1000 == 999:
            print("  step %d, %.0f s" % (step + 1, time.time() - t0), flush=True)
    print("FUSED_STRESS large %dx%d shape %d store_h %d: one-pass launches=%d (Ez compared after every launch, "
          "all 9 arrays every 50) mismatches=%d, %.0f s" % (npx, npy, shape, store_h, step + 1, mism,
                                                            time.time() - t0), flush=True)
    for eng in engines:
        eng.close()
    return mism


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "small"
    if mode == "small":
        a = sys.argv[2:]
        npx, npy, band = (int(x) for x in (a[0:3] if len(a) >= 3 else (300, 700, 48)))
        shapes = [int(x) for x in a[3].split(",")] if len(a) > 3 else [20, 23, 22, 24]
        reps = int(a[4]) if len(a) > 4 else 30
        reset_every = int(a[5]) if len(a) > 5 else 25
        sys.exit(1 if small(npx, npy, band, shapes, reps, reset_every) else 0)
    a = sys.argv[2:]
    sys.exit(1 if large(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]) if len(a) > 4 else 0) else 0)

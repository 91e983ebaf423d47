"""ONE process, several y-slab engines tied together with b200fdtd_peer_attach_engine (direct
stores into the neighbour's ghost columns + device flags, no NCCL, no host synchronisation per
step) must reproduce the single-engine run: fields bit for bit (tolerance forms: 1e-12), far field
to summation-order noise.  The slabs sit on one device or are dealt round-robin over the visible
devices, so a one-GPU box exercises the whole protocol -- including a MIDDLE slab with both
neighbours attached, which is what ranks 1..N-2 of an 8-GPU run execute.

    python scripts/peer_local_check.py SOLVER NPX NPY STEPS WORLD FORM [same|spread] [model|random]
    FORM: exact | unit | fused | lean | leanfused
    random: instead of the Mie cylinder from rest, a random state + random permittivity (every cell
    non-zero from the first step; the columns the ghost columns mirror start at zero, like the ghosts)
"""
import ctypes as C
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # one hardware queue per stream: a spinning
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # flag wait never blocks a neighbour
solver, npx, npy, steps, world, form = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                        int(sys.argv[5]), sys.argv[6])
placement = sys.argv[7] if len(sys.argv) > 7 else "same"
start = sys.argv[8] if len(sys.argv) > 8 else "model"
if form in ("lean", "leanfused"):
    os.environ["B200FDTD_LEAN_INTERIOR"] = "1"
if form in ("fused", "leanfused"):
    os.environ["B200FDTD_FUSED"] = "1"
if form == "unit":
    os.environ["B200FDTD_UNIT_SPLIT"] = "1"
import numpy as np

from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun

L = B.lib()
n_dev = B.device_count()
model, angle, hu = "MIE_CYLINDER", 20, 20


def random_case(world_cut):
    kind = 2 if solver == "TM_UPML_2D" else 3
    rng = np.random.default_rng(99)
    eps = [np.where(rng.random((npx, npy)) < 0.67, 1.0, 1.5 + rng.random((npx, npy))) for _ in range(1 if kind == 2 else 2)]
    if npy >= 1024:                     # wide grids: material in the middle third of the rows only, so that whole
        for e in eps:                   # 256-column tile rows are vacuum (the one-pass step keeps no E arrays there)
            e[:npx // 3, :] = 1.0
            e[2 * npx // 3:, :] = 1.0
    state = [rng.standard_normal((npx, npy)) + 1j * rng.standard_normal((npx, npy)) for _ in range(9)]
    for h, b in (((3, 5), (6, 8)) if kind == 2 else ((6, 8),)):
        state[h] = (state[b].real / B.MU_0_S) + 1j * (state[b].imag / B.MU_0_S)
    from mpifdtd_b200.slab import split_columns
    for arr in state:
        for r in range(1, world_cut):
            j0 = split_columns(npy, world_cut, r)[0]
            arr[:, j0 - 2:j0 + 2] = 0
    return eps, state


CASE = random_case(world) if start == "random" else None
WIDE = npy >= 1024


def run(world):
    runs = [SlabRun("NO_MODEL" if CASE else model, solver, npx, npy, steps, rank=r, world=world,
                    device=(r % n_dev if placement == "spread" else 0), h_u_nm=hu, angle_deg=angle,
                    n_bins="full" if (CASE or WIDE) else None)
            for r in range(world)]
    if CASE:
        for r in runs:
            for slot, e in enumerate(CASE[0]):
                r.engine.set_eps(slot, e)
            for slot in range(9):
                r.engine.set_field(slot, CASE[1][slot])
    for lo, hi in zip(runs[:-1], runs[1:]):
        lo.engine.peer_attach_engine(1, hi.engine)
        hi.engine.peer_attach_engine(0, lo.engine)
    args = B.StepArgs()
    L.field_reset()
    for _ in range(steps):
        L.mpifdtd_upml_step_args(runs[0].kind, 0, C.byref(args))
        for r in runs:                       # asynchronous: the slabs order themselves with the device flags
            r.engine.step(args)
        L.field_nextStep()
    for r in runs:
        r.engine.sync()
    fields = [np.concatenate([r.gather_field(s) for r in runs], axis=1) for s in range(9)]
    digests = [sum(r.engine.digest(s) for r in runs) & (2**64 - 1) for s in range(9)]
    for r in runs:
        r.project()
    for r in runs[1:]:
        runs[0].engine.add_uw(r.engine)
    if CASE or WIDE:  # a run this short (or a surface this wide) leaves the far field's own bins empty: compare the whole U/W block
        far = np.stack([runs[0].engine.uw(s) for s in range(3)])
    else:
        far = np.zeros((321, 360))
        L.mpifdtd_upml_far_field(runs[0].engine.h, runs[0].kind, 0, far.ctypes.data)
    forms = [r.engine.step_form() for r in runs]
    for r in runs:
        r.close()
    return fields, digests, far, forms


want, want_d, want_far, _ = run(1)
got, got_d, got_far, forms = run(world)
ok = True
exact = form in ("exact", "unit", "fused")
assert np.abs(want[0]).max() > 0
for s in range(9):
    if exact:
        same = np.array_equal(got[s].view(np.float64), want[s].view(np.float64)) and got_d[s] == want_d[s]
        print("slot", s, "bit-identical:", same)
    else:
        if s in (1, 4, 7):       # M / J: stale outside the frame in the tolerance forms
            continue
        err = np.abs(got[s] - want[s]).max() / np.abs(want[s]).max()
        same = err <= 1e-12
        print("slot", s, "rel err", err)
    ok &= bool(same)
err = np.abs(got_far - want_far).max() / np.abs(want_far).max() if np.abs(want_far).max() > 0 else float("inf")
print("far field rel err vs single engine:", err, "step forms of the slabs:", forms)
ok &= err < (1e-12 if exact else 1e-10)
print("PEER_LOCAL_CHECK", solver, form, "world", world, placement, start, "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)

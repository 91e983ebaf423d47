"""Every form of a UPML step must produce identical bits:

  * two-kernel step that stores H (the literal restatement of the reference's passes),
  * two-kernel step that does NOT store H and forms it as B/mu0 in the E phase (default),
  * the one-pass kernel (fused_kernels.cu: H and E phase of a row band in one march, operands
    staged by TMA bulk copies behind an mbarrier ring, producer warp + consumer warps), with and
    without H stores, in every launch shape -- for TM (fdtdTM_upml.c:155-219) and TE
    (fdtdTE_upml.c:252-314).

All evaluate the same expressions in the same order (-fmad=false), so after any number of
steps from any state each of the nine arrays and the NTFF history must agree bit for bit
-- including ragged strips (columns not a multiple of 32), bands of any height and strips
narrower than a warp.  The lean (tolerance) form has the same property between ITS two-kernel
and one-pass realisations."""
import ctypes as C

import numpy as np
import pytest

from helpers import bit_equal
from mpifdtd_b200 import binding as B

pytestmark = pytest.mark.gpu

H_OF_B = {2: ((3, 5), (6, 8)), 3: ((6, 8),)}          # slots (H, B) tied by H == B/mu0 after any H phase


def make_engine(L, npx, npy, steps, eps, fused, store_h=0, band=None, j0=0, nj=None, shape=20, kind=2, lean=0):
    eng = B.Engine(kind, npx, npy, 10, j0=j0, nj=nj)
    ti, tj = np.empty((6, npx)), np.empty((6, npy))
    L.mpifdtd_upml_tables(kind, ti.ctypes.data, tj.ctypes.data)
    eng.set_tables(ti, tj)
    for slot, e in enumerate(eps if isinstance(eps, (list, tuple)) else [eps]):
        eng.set_eps(slot, e)
    box = L.field_getNTFFInfo()
    n_local = L.mpifdtd_ntff_local_count(C.byref(box), eng.j0, eng.nj)
    ptr = L.mpifdtd_ntff_time_shift(C.byref(box), 360, 0.0 if kind == 2 else 0.5, eng.j0, eng.nj)
    plan = B.NtffPlan(box.top, box.bottom, box.left, box.right,
                      L.mpifdtd_ntff_point_count(C.byref(box)), n_local, steps, steps, 360,
                      box.arraySize, ptr)
    B.check(L.b200fdtd_set_ntff_plan(eng.h, C.byref(plan)), "plan")
    L.free(ptr)
    eng.n_bins = steps
    eng.set_option(B.OPT_LEAN_INTERIOR, lean)
    eng.set_option(B.OPT_FUSED, fused)
    eng.set_option(B.OPT_STORE_H, store_h)
    if fused and band:
        eng.set_option(B.OPT_BAND_ROWS, band)
    if fused:
        eng.set_option(B.OPT_FUSED_SHAPE, shape)
    return eng


def random_case(npx, npy, kind):
    rng = np.random.default_rng(npx * 1000 + npy + kind)
    eps = [np.where(rng.random((npx, npy)) < 0.5, 1.0, 1.0 + 2.0 * rng.random((npx, npy)))
           for _ in range(1 if kind == 2 else 2)]
    state = [rng.standard_normal((npx, npy)) + 1j * rng.standard_normal((npx, npy)) for _ in range(9)]
    # H arrays must satisfy the solver's invariant H == B/mu0 (true after any H phase)
    mu0 = B.MU_0_S
    for h, b in H_OF_B[kind]:
        state[h] = (state[b].real / mu0) + 1j * (state[b].imag / mu0)
    return eps, state


def run_and_compare(L, kind, engines, state, steps):
    for eng in engines:
        for slot in range(9):
            eng.set_field(slot, state[slot])
    args = B.StepArgs()
    L.field_reset()
    for _ in range(steps):
        L.mpifdtd_upml_step_args(kind, 1, C.byref(args))       # pulse + point source both on
        for eng in engines:
            eng.step(args)
        L.field_nextStep()
    ref = [engines[0].get_field(s) for s in range(9)]
    assert np.abs(ref[0]).max() > 0 and np.all(np.isfinite(ref[0].view(np.float64)))
    for which, eng in enumerate(engines[1:], 1):
        for slot in range(9):
            got = eng.get_field(slot)
            if not bit_equal(got.view(np.float64), ref[slot].view(np.float64)):      # say where, for the post-mortem
                ii, jj = np.nonzero(got != ref[slot])
                raise AssertionError("engine %d slot %d: %d cells differ, rows %d..%d, columns %d..%d"
                                     % (which, slot, len(ii), ii.min(), ii.max(), jj.min(), jj.max()))
    for eng in engines:
        eng.project()
    for slot in range(3):
        want = engines[0].uw(slot)
        for eng in engines[1:]:
            assert bit_equal(eng.uw(slot).view(np.float64), want.view(np.float64)), slot
    assert engines[1].launches() > 0
    for eng in engines:
        eng.close()


CASES = [(70, 96, 256), (45, 47, 7), (131, 200, 64), (300, 41, 33), (64, 1030, 1), (257, 66, 256), (300, 700, 48)]


@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("npx,npy,band", CASES)
def test_fused_step_is_bit_identical_to_two_kernel_step(plugin_lib, npx, npy, band, kind):
    L = plugin_lib
    steps = 6
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
    eps, state = random_case(npx, npy, kind)
    mk = lambda *a, **kw: make_engine(L, npx, npy, steps, eps, *a, kind=kind, **kw)
    engines = [mk(0, store_h=1), mk(0, store_h=0),
               mk(1, store_h=0, band=band, shape=20),      # 256 columns x 4 row buffers: the default
               mk(1, store_h=1, band=band, shape=20),
               mk(1, store_h=1, band=band, shape=23),      # 128 columns x 8 row buffers
               mk(1, store_h=0, band=band, shape=23),
               mk(1, store_h=1, band=band, shape=24),      # 256 columns x 3
               mk(1, store_h=0, band=band, shape=21),      # 256 columns x 6
               mk(1, store_h=0, band=band, shape=22)]      # 512 columns x 3
    run_and_compare(L, kind, engines, state, steps)


@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("npx,npy,band", CASES)
def test_lean_one_pass_is_bit_identical_to_lean_two_kernel_step(plugin_lib, npx, npy, band, kind):
    """The tolerance form: cells of the frame-free rectangle advance B / D directly, in two
    kernels (upml_kernels.cu) or in one pass -- same expressions, same bits; M / J of those cells
    are touched by neither."""
    L = plugin_lib
    steps = 6
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
    eps, state = random_case(npx, npy, kind)
    mk = lambda *a, **kw: make_engine(L, npx, npy, steps, eps, *a, kind=kind, lean=1, **kw)
    # (the two-kernel lean step that STORES H forms curl H from the stored quotients instead of
    # RN(1/mu0) * curl B -- another rounding, same tolerance -- so the reference here is the default,
    # H-not-stored one; the one-pass kernel uses the B form whether or not it also stores H)
    engines = [mk(0, store_h=0),
               mk(1, store_h=0, band=band, shape=20), mk(1, store_h=1, band=band, shape=20),
               mk(1, store_h=0, band=band, shape=23), mk(1, store_h=1, band=band, shape=21),
               mk(1, store_h=0, band=band, shape=22), mk(1, store_h=1, band=band, shape=24)]
    assert engines[0].step_form() == 2 and engines[1].step_form() == 4
    run_and_compare(L, kind, engines, state, steps)


def structured_case(npx, npy, kind):
    """vacuum with a dielectric block and a few single cells: most rows of most strips have eps == 1
    everywhere, some do not; random state"""
    rng = np.random.default_rng(7 * npx + npy + kind)
    eps = []
    for m in range(1 if kind == 2 else 2):
        e = np.ones((npx, npy))
        e[npx // 3:npx // 3 + 9, npy // 2 - 40:npy // 2 + 200 + 30 * m] = 2.56
        for _ in range(6):
            e[rng.integers(12, npx - 12), rng.integers(12, npy - 12)] = 1.0 + rng.random()
        eps.append(e)
    state = [rng.standard_normal((npx, npy)) + 1j * rng.standard_normal((npx, npy)) for _ in range(9)]
    mu0 = B.MU_0_S
    for h, b in H_OF_B[kind]:
        state[h] = (state[b].real / mu0) + 1j * (state[b].imag / mu0)
    return eps, state


@pytest.mark.parametrize("lean", [0, 1])
@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("npx,npy,band", [(150, 1100, 32), (97, 790, 7), (140, 1560, 63), (60, 530, 64)])
def test_vacuum_row_strips_keep_no_e_arrays_and_change_no_bit(plugin_lib, npx, npy, band, kind, lean):
    """B200FDTD_OPT_DERIVED_E: in a row of a CTA tile whose cells all have eps == 1 the E arrays hold the
    bits of D (E = D/1.0), so the pass reads D for the old E, stores no E and stages no eps there.
    From a random state (first step: E arrays everywhere, since E == D/eps is not given yet), with a
    getter in mid-run (refresh of the E arrays, then derived again), a changed permittivity map and a
    restart: all nine arrays and the NTFF history bit-identical to the two-kernel step and to the pass
    with the option off.  band 64 does not fit the 64-bit row masks: the option is then inert."""
    L = plugin_lib
    steps = 14
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
    eps, state = structured_case(npx, npy, kind)
    mk = lambda *a, **kw: make_engine(L, npx, npy, steps, eps, *a, kind=kind, lean=lean, **kw)
    engines = [mk(0), mk(1, band=band, shape=20), mk(1, band=band, shape=20), mk(1, band=band, shape=23),
               mk(1, band=band, shape=22 if kind == 2 else 21)]
    engines[1].set_option(B.OPT_DERIVED_E, 0)
    assert engines[1].vacuum_cells() == 0
    for eng in engines[2:]:
        n_vac = eng.vacuum_cells()
        assert (n_vac == 0) if band > 63 else (0 < n_vac < npx * npy), n_vac

    def compare(what):
        ref = [engines[0].get_field(s) for s in range(9)]
        assert np.abs(ref[0]).max() > 0 and np.all(np.isfinite(ref[0].view(np.float64)))
        for which, eng in enumerate(engines[1:], 1):
            for slot in range(9):
                got = eng.get_field(slot)
                if not bit_equal(got.view(np.float64), ref[slot].view(np.float64)):
                    ii, jj = np.nonzero(got != ref[slot])
                    raise AssertionError("%s: engine %d slot %d: %d cells differ, rows %d..%d, columns %d..%d"
                                         % (what, which, slot, len(ii), ii.min(), ii.max(), jj.min(), jj.max()))

    def advance(n):
        args = B.StepArgs()
        for _ in range(n):
            L.mpifdtd_upml_step_args(kind, 0, C.byref(args))       # the pulse only
            for eng in engines:
                eng.step(args)
            L.field_nextStep()

    for eng in engines:
        for slot in range(9):
            eng.set_field(slot, state[slot])
    L.field_reset()
    advance(5)
    compare("after 5 steps")
    advance(4)
    compare("after 9 steps")
    eps2 = [e.copy() for e in eps]                              # material where there was vacuum: new masks
    for e in eps2:
        e[npx // 2, 20:npy - 20:37] = 1.44
    for eng in engines:
        for slot, e in enumerate(eps2):
            eng.set_eps(slot, e)
    advance(5)
    compare("after 14 steps, new permittivity")
    for eng in engines:
        eng.project()
    for slot in range(3):
        want = engines[0].uw(slot)
        for eng in engines[1:]:
            assert bit_equal(eng.uw(slot).view(np.float64), want.view(np.float64)), slot
    for eng in engines:                                         # from rest: E == D == 0 is consistent at once
        eng.zero()
        eng.set_field(1 if kind == 2 else 4, state[1])          # ... but a setter call is not
    L.field_reset()
    advance(3)
    compare("after a restart")
    for eng in engines:
        eng.close()


def test_one_pass_replay_from_a_set_state(plugin_lib):
    """b200fdtd_run_steps with the one-pass step after b200fdtd_set_field: the first step runs outside the
    graph (E arrays everywhere), the captured ones keep no E in the vacuum row-strips; same bits as
    stepping one by one with the option off."""
    L = plugin_lib
    npx, npy, steps, kind = 120, 800, 40, 2
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 30, steps))
    eps, state = structured_case(npx, npy, kind)
    plain = make_engine(L, npx, npy, steps, eps, 1, kind=kind)
    plain.set_option(B.OPT_DERIVED_E, 0)
    replay = make_engine(L, npx, npy, steps, eps, 1, kind=kind)
    for eng in (plain, replay):
        for slot in range(9):
            eng.set_field(slot, state[slot])
    L.field_reset()
    args = B.StepArgs()
    for _ in range(steps):
        L.mpifdtd_upml_step_args(kind, 0, C.byref(args))
        plain.step(args)
        L.field_nextStep()
    src = (C.c_char * L.b200fdtd_struct_size(5))()
    L.field_reset()
    L.mpifdtd_upml_batch_source.argtypes = [C.c_int, C.c_double, C.c_void_p]
    L.mpifdtd_upml_batch_source(kind, 30.0, src)
    B.check(L.b200fdtd_set_batch_sources(replay.h, src), "set_batch_sources")
    L.b200fdtd_run_steps.argtypes = [C.c_void_p, C.c_double, C.c_int32]
    B.check(L.b200fdtd_run_steps(replay.h, 0.0, steps), "run_steps")
    assert replay.vacuum_cells() > 0
    for slot in range(9):
        assert bit_equal(replay.get_field(slot).view(np.float64), plain.get_field(slot).view(np.float64)), slot
    plain.close(); replay.close()


def test_getter_between_two_replays_of_the_same_graph(plugin_lib, in_tmp_cwd, monkeypatch):
    """The plugin on a grid of >= 2^22 cells: one-pass step, update() calls replayed from a 128-step CUDA graph.
    A getter after the first replay brings the E arrays up to date; the second replay of the SAME graph (no
    launch function runs) leaves them behind D again in the vacuum row-strips, and the next getter must know."""
    npx, npy, steps = 2048, 2100, 300
    monkeypatch.setenv("MPIFDTD_DEFER_CHUNK", "128")
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B200FDTD_DERIVED_E", mode)
        gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", npx, npy, steps=steps, h_u_nm=10, angle_deg=20)
        snaps = []
        for n in (128, 128, 44):
            gpu.step(n)
            snaps.append(gpu.field("Ez"))
        snaps += [gpu.any_field(s) for s in range(9)]
        res[mode] = snaps
        gpu.finish()
    assert np.abs(res["0"][1]).max() > 0 and not np.array_equal(res["0"][0], res["0"][1])
    for n, (a, b) in enumerate(zip(res["1"], res["0"])):
        assert bit_equal(a, b), n


def test_launches_per_step(plugin_lib, in_tmp_cwd, monkeypatch):
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", 96, steps=10, h_u_nm=20)
    n0 = gpu.launches()
    gpu.step(1)
    assert gpu.launches() - n0 == 4          # H phase, E phase + source, NTFF sample, device clock tick
    gpu.finish()
    monkeypatch.setenv("MPIFDTD_DEFER_STEPS", "0")      # every update() handed over on its own: no clock
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", 96, steps=10, h_u_nm=20)
    n0 = gpu.launches()
    gpu.step(1)
    assert gpu.launches() - n0 == 3
    gpu.finish()


def test_permittivity_division_is_exactly_ieee(plugin_lib):
    """div_eps(z, eps) -- correctly rounded reciprocal + two Markstein corrections, one reciprocal
    per cell -- must equal z / eps bit for bit: 2^31 operand pairs with a fresh divisor each."""
    bad = C.c_uint64(123)
    B.check(plugin_lib.b200fdtd_selftest_division(0.0, 1 << 31, C.byref(bad)), "selftest")
    assert bad.value == 0


@pytest.mark.parametrize("divisor", [B.MU_0_S, 2.56, 1.0 / 3.0, 1.5625, 15.069924])
def test_constant_division_shortcut_is_exactly_ieee(plugin_lib, divisor):
    """div_const(x, d) must equal x / d bit for bit (upml_common.cuh): 2^31 operands per
    divisor, half of them raw 64-bit patterns (all exponents, subnormals, inf, nan)."""
    bad = C.c_uint64(123)
    B.check(plugin_lib.b200fdtd_selftest_division(divisor, 1 << 31, C.byref(bad)), "selftest")
    assert bad.value == 0


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_one_pass_step_is_the_default_on_large_grids(plugin_lib, in_tmp_cwd, monkeypatch, solver):
    """auto (default): grids of >= 2^22 updated cells take the TMA-staged one-pass step -- through
    b200fdtd_step and through the plugin's deferred multi-step replay (device clock, CUDA graph) --
    and produce the bits of the two-kernel step on all nine arrays and the NTFF history."""
    npx, npy, steps = 2048, 2100, 24
    res = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("B200FDTD_FUSED", mode)
        gpu = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=10, angle_deg=20)
        form = C.c_int32(-1)
        B.check(gpu.L.b200fdtd_get_step_form(gpu.engine_handle(), C.byref(form)), "get_step_form")
        assert form.value == (3 if mode == "2" else 1)
        n0 = gpu.launches()
        gpu.run()
        gpu.sync()
        per_step = (gpu.launches() - n0) / steps
        # pre-pass x2 + one pass | (interior + frame) x2; + sample + clock (+ once, the three kernels that build the
        # row masks of the vacuum row-strips)
        assert per_step == (5 + 3 / steps if mode == "2" else 6)
        res[mode] = [gpu.any_field(s) for s in range(9)] + [gpu.ntff_uw(s, project=(s == 0)) for s in range(3)]
        gpu.finish()
    assert np.abs(res["0"][0]).max() > 0
    for n, (a, b) in enumerate(zip(res["2"], res["0"])):
        assert bit_equal(a, b), n
    # small grids keep one kernel per phase
    monkeypatch.setenv("B200FDTD_FUSED", "2")
    small = B.Plugin("MIE_CYLINDER", solver, 256, 256, steps=8, h_u_nm=10)
    B.check(small.L.b200fdtd_get_step_form(small.engine_handle(), C.byref(form)), "get_step_form")
    assert form.value == 0
    small.finish()

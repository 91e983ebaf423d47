"""Pins oracle/fdtd_oracle.c and oracle/split_oracle.c (the plain-C restatements) to the unmodified reference:
golden vectors recorded from the reference (tests/golden/*.npz) and, when present,
the reference library itself.  On the machine that generated the goldens the match
is bit-exact; the asserted tolerance leaves room for a different host libm."""
import os

import numpy as np
import pytest

from helpers import ANGLE_ROWS, GOLDEN, LAMBDA_ROWS, bit_equal, golden, rel_err

RUNS = ["mie_tm_upml_88x96", "mie_te_upml_88x96", "zigzag_tm_upml_72x120_a30", "layer_te_upml_80x110_a45"]


def make_oracle(oracle, g):
    npx, npy, hu, steps, angle, solver, _model = (int(v) for v in g["meta"])
    if solver == 2:
        sim = oracle.OracleSim(oracle.TM, npx, npy, steps, g["EPS_EZ"], h_u_nm=hu, angle_deg=angle)
    else:
        sim = oracle.OracleSim(oracle.TE, npx, npy, steps, g["EPS_EX"], g["EPS_EY"], h_u_nm=hu, angle_deg=angle)
    return sim, steps, solver


@pytest.mark.parametrize("name", RUNS)
def test_oracle_matches_reference_run(oracle, name):
    g = golden(name + ".npz")
    sim, steps, solver = make_oracle(oracle, g)
    coef_names = oracle.TM_COEFS if solver == 2 else oracle.TE_COEFS
    for c in coef_names:
        if c in g.files:
            assert bit_equal(sim.coef(c), g[c]), c
    half = steps // 2
    sim.step(half)
    for key in [k for k in g.files if k.startswith("mid_")]:
        assert rel_err(sim.field(key[4:]), g[key]) <= 1e-13, key
    sim.step(steps - half)
    for key in [k for k in g.files if k.startswith("end_")]:
        assert rel_err(sim.field(key[4:]), g[key]) <= 1e-13, key
    uw_order = ["Ux", "Uy", "Wz"] if solver == 2 else ["Wx", "Wy", "Uz"]
    for key in [k for k in g.files if k.startswith("uw_")]:
        mine = sim.uw(uw_order.index(key[3:]))[ANGLE_ROWS, :steps]
        assert rel_err(mine, g[key]) <= 1e-13, key
    assert rel_err(sim.far_field()[LAMBDA_ROWS, :], g["far_field_rows"]) <= 1e-12
    sim.close()


def test_oracle_fft_matches_reference_cfft(oracle):
    g = golden("cfft_256.npz")
    assert rel_err(oracle.fft(g["x"]), g["y"]) <= 1e-15
    # convention check (cfft.c:131-141): unnormalised, positive exponent = N * ifft
    assert rel_err(oracle.fft(g["x"]), 256 * np.fft.ifft(g["x"])) <= 1e-12


def test_oracle_vs_live_reference(oracle):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    n, steps = 96, 260
    ref = reflib.RefSim("MIE_CYLINDER", "TM_UPML_2D", n, steps=steps, h_u_nm=20)
    sim = oracle.OracleSim(oracle.TM, n, n, steps, ref.coef("EPS_EZ"), h_u_nm=20)
    ref.run()
    sim.step(steps)
    for f in ("Ez", "Hx", "Hy", "Jz", "Dz", "Mx", "Bx", "My", "By"):
        assert bit_equal(sim.field(f).view(np.float64), ref.field(f).view(np.float64)), f
    for slot, name in enumerate(("Ux", "Uy", "Wz")):
        assert bit_equal(sim.uw(slot).view(np.float64), ref.ntff_uw(name).view(np.float64)), name
    assert bit_equal(sim.far_field(), ref.finish())
    sim.close()


# ---------------------------------------------------------------- split-field solvers (ids 0, 1, 6, 7)
EPS_NAMES = {0: ["EPS_EZ", "EPS_HX", "EPS_HY"], 1: ["EPS_EX", "EPS_EY", "EPS_HZ"]}
EPS_NAMES[6], EPS_NAMES[7] = EPS_NAMES[0], EPS_NAMES[1]


@pytest.mark.parametrize("kind,model,angle", [(0, "MIE_CYLINDER", 0), (1, "LAYER", 15), (6, "ZIGZAG", 30),
                                              (7, "MIE_CYLINDER", 15), (7, "LAYER", 0), (6, "MORPHO_SCALE", 0)])
def test_split_oracle_is_bit_exact_vs_live_reference(oracle, kind, model, angle):
    """oracle/split_oracle.c against the unmodified reference: the eight coefficient arrays and
    all five fields after 260 steps, bit for bit (same host libm, same expression order)."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    npx, npy, steps, lam = 96, 110, 260, 633
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, lambda_nm=lam, angle_deg=angle)
    eps = [ref.coef(n) for n in EPS_NAMES[kind]]
    cpu = oracle.SplitOracleSim(kind, npx, npy, eps, lambda_nm=lam, angle_deg=angle)
    inner = (slice(1, -1), slice(1, -1)) if kind == 7 else (slice(None), slice(None))
    for n in oracle.SPLIT_COEFS[kind]:
        assert bit_equal(np.ascontiguousarray(cpu.coef(n)[inner]), np.ascontiguousarray(ref.coef(n)[inner])), n
    ref.run()
    cpu.step(steps)
    assert np.abs(ref.field(oracle.SPLIT_FIELDS[kind][0])).max() > 1e-3
    for f in oracle.SPLIT_FIELDS[kind]:
        assert bit_equal(cpu.field(f), ref.field(f)), f
    cpu.close()


@pytest.mark.parametrize("kind", [0, 1, 6, 7])
def test_split_oracle_vs_reference_recorded_snapshots(oracle, kind):
    g = np.load(os.path.join(GOLDEN, "split_kind%d.npz" % kind))
    npx, npy, hu, steps, lam, angle = (int(v) for v in g["meta"][:6])
    cpu = oracle.SplitOracleSim(kind, npx, npy, [g[n] for n in EPS_NAMES[kind]], h_u_nm=hu, lambda_nm=lam,
                                angle_deg=angle)
    cpu.step(steps // 2)
    for f in oracle.SPLIT_FIELDS[kind]:
        assert rel_err(cpu.field(f), g["mid_" + f]) <= 1e-13, ("mid", f)
    cpu.step(steps - steps // 2)
    for f in oracle.SPLIT_FIELDS[kind]:
        assert rel_err(cpu.field(f), g["end_" + f]) <= 1e-13, ("end", f)
    cpu.close()


MPI_FIELDS = {4: ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"],
              5: ["Ex", "Jx", "Dx", "Ey", "Jy", "Dy", "Hz", "Mz", "Bz"]}


@pytest.mark.parametrize("kind,model,angle", [(4, "MIE_CYLINDER", 25), (5, "MIE_CYLINDER", 25), (4, "ZIGZAG", 0),
                                              (5, "LAYER", 60)])
def test_mpi_variant_restatement_vs_live_reference(oracle, plugin_lib, kind, model, angle):
    """oracle_step_mpi (solver ids 4 / 5 at one rank: E first, CW source, every cell against a zero
    ghost ring, then the variants' own ntff()) against the unmodified reference: all nine arrays, the
    U/W accumulators, and the coefficient arrays it shares with the serial restatement against the
    reference's own (N+2) x (N+2) ones."""
    from oracle import reflib
    from mpifdtd_b200 import binding as B
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    npx, npy, steps = 90, 104, 240
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, angle_deg=angle)
    sub = (npx + 2) * (npy + 2)
    grab = lambda name, real=False: (ref.darray(name, sub) if real else ref.carray(name, sub)).reshape(npx + 2, npy + 2)[1:-1, 1:-1]
    if kind == 4:
        sim = oracle.OracleSim(oracle.TM, npx, npy, steps, grab("EPS_EZ", True), angle_deg=angle)
        coefs = oracle.TM_COEFS
    else:
        sim = oracle.OracleSim(oracle.TE, npx, npy, steps, grab("EPS_EX", True), grab("EPS_EY", True), angle_deg=angle)
        coefs = oracle.TE_COEFS
    for c in coefs:
        assert bit_equal(sim.coef(c), grab(c, True)), c
    for chunk in (steps // 3, steps - steps // 3):
        ref.step(chunk)
        sim.step_mpi(chunk)
        for slot, f in enumerate(MPI_FIELDS[kind]):
            want = grab(f)
            assert rel_err(sim.field(slot), want) <= 1e-13, (f, chunk)
    assert np.abs(grab(MPI_FIELDS[kind][0])).max() > 1e-3
    # the variants' own ntff() (row a15): all arraySize bins of all 360 directions, spills included
    for slot, name in enumerate(["Ux", "Uy", "Wz"] if kind == 4 else ["Wx", "Wy", "Uz"]):
        want = ref.ntff_uw(name)
        assert np.abs(want).max() > 0, name
        assert rel_err(sim.uw(slot), want) <= 1e-13, name
    os.chdir(cwd)          # (the reference's finish() for these ids calls MPI_Finalize; not needed)
    sim.close()


@pytest.mark.parametrize("kind,model,form,hook", [
    (2, "MIE_CYLINDER", "CW", "refhook_tm_upml_update_cw"),
    (2, "NO_MODEL", "PLANE", "refhook_tm_upml_update_plane_wave"),
    (4, "NO_MODEL", "PLANE", "refhook_mpi_tm_upml_update_plane_wave"),
    (4, "ZIGZAG", "PLANE", "refhook_mpi_tm_upml_update_plane_wave")])
def test_optional_source_restatements_vs_live_reference(oracle, kind, model, form, hook):
    """Row a10's opt-in sources (field_scatteredWave, planeWave) in the C restatement against the
    reference's own static functions as oracle/refbuild/wrap_*.c compose them."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    npx, npy, steps = (128, 128, 200) if model == "NO_MODEL" else (90, 104, 240)
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, angle_deg=20)
    ring = 1 if kind == 4 else 0
    sub = (npx + 2 * ring) * (npy + 2 * ring)

    def grab(name, real=False):
        a = (ref.darray(name, sub) if real else ref.carray(name, sub)).reshape(npx + 2 * ring, npy + 2 * ring)
        return a[1:-1, 1:-1] if ring else a

    sim = oracle.OracleSim(oracle.TM, npx, npy, steps, grab("EPS_EZ", True), angle_deg=20, source_form=form)
    ref.step_fn(hook, steps)
    if kind == 4:
        sim.step_mpi(steps)
    else:
        sim.step(steps)
    assert np.abs(grab("Ez")).max() > 1e-3
    for slot, f in ((0, "Ez"), (3, "Hx"), (6, "Hy")):
        assert rel_err(sim.field(slot), grab(f)) <= 1e-13, f
    for slot, name in enumerate(("Ux", "Uy", "Wz")):
        want = ref.ntff_uw(name)
        assert rel_err(sim.uw(slot), want) <= 1e-13, name
    os.chdir(cwd)
    sim.close()


@pytest.mark.parametrize("name", ["mpi_kind4_mie", "mpi_kind5_mie", "mpi_kind4_zigzag_plane"])
def test_mpi_variant_restatement_matches_reference_recording(oracle, name):
    """oracle_step_mpi against arrays recorded from the unmodified reference (ids 4 / 5 at one rank,
    tests/golden/make_golden.py mpi): pins the restatement where libref.so is absent."""
    g = golden(name + ".npz")
    npx, npy, hu, steps, angle, kind, _model, plane = (int(v) for v in g["meta"])
    if kind == 4:
        sim = oracle.OracleSim(oracle.TM, npx, npy, steps, g["EPS_EZ"], h_u_nm=hu, angle_deg=angle,
                               source_form="PLANE" if plane else 0)
        fields, uws = {"Ez": 0, "Hx": 3, "Hy": 6}, ["Ux", "Uy", "Wz"]
    else:
        sim = oracle.OracleSim(oracle.TE, npx, npy, steps, g["EPS_EX"], g["EPS_EY"], h_u_nm=hu, angle_deg=angle)
        fields, uws = {"Ex": 0, "Ey": 3, "Hz": 6}, ["Wx", "Wy", "Uz"]
    sim.step_mpi(steps // 2)
    for f, slot in fields.items():
        assert rel_err(sim.field(slot), g["mid_" + f]) <= 1e-13, f
    sim.step_mpi(steps - steps // 2)
    for f, slot in fields.items():
        assert rel_err(sim.field(slot), g["end_" + f]) <= 1e-13, f
    for slot, u in enumerate(uws):
        assert rel_err(sim.uw(slot)[ANGLE_ROWS, :], g["uw_" + u]) <= 1e-13, u
    sim.close()

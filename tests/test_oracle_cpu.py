"""Pins oracle/fdtd_oracle.c (the plain-C restatement) to the unmodified reference:
golden vectors recorded from the reference (tests/golden/*.npz) and, when present,
the reference library itself.  On the machine that generated the goldens the match
is bit-exact; the asserted tolerance leaves room for a different host libm."""
import numpy as np
import pytest

from helpers import ANGLE_ROWS, LAMBDA_ROWS, bit_equal, golden, rel_err

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

"""The pipelined step (one persistent kernel per time step: H phase of row band k+1 side by
side with the E phase of band k, so the E phase finds B in L2) runs the same per-cell code as
the two-kernel step and must therefore be bit-identical to it -- every array, NTFF history
included -- for TM and TE, double and single precision, any band height, ragged grids, and
many steps (the inter-CTA ordering is what is under test)."""
import numpy as np
import pytest

from helpers import bit_equal
from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun

pytestmark = pytest.mark.gpu


def run(model, solver, npx, npy, steps, pipelined, band=None, precision="f64", angle=20):
    r = SlabRun(model, solver, npx, npy, steps, angle_deg=angle, precision=precision, h_u_nm=20)
    r.engine.set_option(B.OPT_PIPELINED, 1 if pipelined else 0)
    if band:
        r.engine.set_option(B.OPT_PIPE_BAND_ROWS, band)
    for _ in range(steps):
        r.step()
    fields = [r.gather_field(s) for s in range(9)]
    r.project()
    uw = [r.engine.uw(s) for s in range(3)]
    launches = r.engine.launches()
    r.close()
    return fields, uw, launches


@pytest.mark.parametrize("solver,model,npx,npy,band,precision", [
    ("TM_UPML_2D", "MIE_CYLINDER", 120, 120, None, "f64"),
    ("TE_UPML_2D", "MIE_CYLINDER", 120, 120, None, "f64"),
    ("TM_UPML_2D", "ZIGZAG", 97, 301, 1, "f64"),
    ("TE_UPML_2D", "LAYER", 301, 97, 3, "f64"),
    ("TM_UPML_2D", "MIE_CYLINDER", 110, 530, 64, "f64"),        # one band holds the whole grid
    ("TM_UPML_2D", "MIE_CYLINDER", 120, 120, 2, "f32"),
    ("TE_UPML_2D", "ZIGZAG", 100, 140, 5, "f32")])
def test_pipelined_step_is_bit_identical(plugin_lib, solver, model, npx, npy, band, precision):
    steps = 540          # the pulse peaks at step 500
    want_f, want_uw, n_two = run(model, solver, npx, npy, steps, False, precision=precision)
    got_f, got_uw, n_pipe = run(model, solver, npx, npy, steps, True, band=band, precision=precision)
    assert np.abs(want_f[0]).max() > 1e-4
    for s in range(9):          # the H arrays come out of the getters' B/mu0 refresh in both forms
        assert bit_equal(got_f[s], want_f[s]), s
    for s in range(3):
        assert bit_equal(got_uw[s], want_uw[s]), s
    assert n_pipe < n_two                      # one step kernel instead of two


def test_pipelined_h_getters_and_large_grid(plugin_lib):
    """4096 x 4096 (many CTAs in flight, thousands of bands) and the derived-H getters."""
    n, steps = 4096, 12
    want_f, _, _ = run("ZIGZAG", "TM_UPML_2D", n, n, steps, False)
    got_f, _, _ = run("ZIGZAG", "TM_UPML_2D", n, n, steps, True)
    for s in range(9):
        assert bit_equal(got_f[s], want_f[s]), s

"""Opt-in "lean interior" form of the UPML solvers (B200FDTD_OPT_LEAN_INTERIOR / env
B200FDTD_LEAN_INTERIOR=1): cells outside the absorbing frame, where every coefficient of
fdtdTM_upml.c:253-271 / fdtdTE_upml.c:384-403 is exactly 1, advance B and D directly and skip
the M / J recurrences that cancel there (168 instead of 264 B per TM cell-update).

Same mathematics, fewer roundings: NOT bit-identical to the reference, so this file holds the
form to the north_star tolerances against the CPU oracle --
  field snapshots <= 1e-12, U/W and far-field table <= 1e-10 (max-abs-diff / max-abs-ref) --
and checks that the lean cells really were taken (M / J stay untouched outside the frame)."""
import numpy as np
import pytest

from helpers import TOL_FARFIELD, TOL_FIELD, rel_err
from mpifdtd_b200 import binding as B
from test_gpu_parity import oracle_for, run_slabs

pytestmark = pytest.mark.gpu

TM_FIELDS = ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"]
TE_FIELDS = ["Ex", "Jx", "Dx", "Ey", "Jy", "Dy", "Hz", "Mz", "Bz"]
AUX = {"TM_UPML_2D": ["Jz", "Mx", "My"], "TE_UPML_2D": ["Jx", "Jy", "Mz"]}


@pytest.fixture(autouse=True)
def _lean(in_tmp_cwd, monkeypatch):
    monkeypatch.setenv("B200FDTD_LEAN_INTERIOR", "1")
    yield
    B.lib().mpifdtd_setPrecision(0)


@pytest.mark.parametrize("solver,names", [("TM_UPML_2D", TM_FIELDS), ("TE_UPML_2D", TE_FIELDS)])
@pytest.mark.parametrize("model,angle", [("MIE_CYLINDER", 0), ("ZIGZAG", 30), ("LAYER", 45)])
def test_lean_interior_vs_oracle(plugin_lib, oracle, solver, names, model, angle):
    npx, npy, steps, hu = 150, 170, 700, 20
    gpu = B.Plugin(model, solver, npx, npy, steps=steps, h_u_nm=hu, angle_deg=angle)
    i_lo, i_hi, j_lo, j_hi = gpu.lean_extent()
    # pml = 10: the frame-free cells are exactly those whose three field positions have sigma == 0
    assert (i_lo, i_hi, j_lo, j_hi) == (10, npx - 11, 10, npy - 11)
    cpu = oracle_for(oracle, gpu, steps, hu=hu, angle=angle)
    gpu.run()
    cpu.step(steps)
    worst = 0.0
    for slot, name in enumerate(names):
        want = cpu.field(slot)
        assert np.abs(want).max() > 0, name
        if name in AUX[solver]:
            # inside the frame the recurrences run as in the reference; outside they are skipped
            mine = gpu.any_field(slot)
            frame = np.ones((npx, npy), dtype=bool)
            frame[i_lo:i_hi + 1, j_lo:j_hi + 1] = False
            assert np.all(mine[~frame] == 0), name
            assert np.abs(mine[frame] - want[frame]).max() <= TOL_FIELD * np.abs(want).max(), name
            continue
        err = rel_err(gpu.any_field(slot), want)
        worst = max(worst, err)
        assert err <= TOL_FIELD, (name, err)
    # tolerance form: differs from the reference's bits, but by rounding only
    assert 0 < worst < 1e-13, worst
    for slot in range(3):
        assert rel_err(gpu.ntff_uw(slot, project=(slot == 0)), cpu.uw(slot)[:, :steps]) <= TOL_FARFIELD, slot
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_lean_off_is_the_bit_exact_default(plugin_lib, solver, monkeypatch):
    """Without the switch nothing changes: extent empty, M / J updated everywhere."""
    monkeypatch.delenv("B200FDTD_LEAN_INTERIOR")
    gpu = B.Plugin("MIE_CYLINDER", solver, 96, 96, steps=200, h_u_nm=20)
    lo_i, hi_i, lo_j, hi_j = gpu.lean_extent()
    assert lo_i > hi_i and lo_j > hi_j
    gpu.run()
    # slots 1, 4, 7: Jz, Mx, My (TM) / Jx, Jy, Mz (TE)
    assert all(np.abs(gpu.any_field(s)[30:60, 30:60]).max() > 0 for s in (1, 4, 7))
    gpu.finish()


@pytest.mark.parametrize("solver", ["MPI_TM_UPML_2D", "MPI_TE_UPML_2D"])
def test_lean_mpi_variants_vs_exact_form(plugin_lib, solver, monkeypatch):
    """Solver ids 4 / 5 (E-first order, CW source, all N x N cells) through the same kernels."""
    n, steps = 120, 400
    names = ("Ez", "Hx", "Hy") if "TM" in solver else ("Ex", "Ey", "Hz")
    lean = B.Plugin("MIE_CYLINDER", solver, n, n, steps=steps, h_u_nm=20)
    assert lean.lean_extent()[0] <= lean.lean_extent()[1]
    lean.run()
    got = {f: lean.field(f) for f in names}
    lean.finish()
    monkeypatch.delenv("B200FDTD_LEAN_INTERIOR")
    exact = B.Plugin("MIE_CYLINDER", solver, n, n, steps=steps, h_u_nm=20)
    exact.run()
    for f in names:
        want = exact.field(f)
        assert np.abs(want).max() > 0
        assert rel_err(got[f], want) <= TOL_FIELD, f
    exact.finish()


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_lean_slab_split_vs_single_engine(plugin_lib, solver):
    """y-slab split in the lean form.  A slab's first owned column stays with the frame kernels (its
    low-side neighbour is a halo column only the H array holds), so those cells round differently
    from the single engine's: equal to rounding, not bit for bit (the default form is, see
    test_gpu_parity.py::test_slab_split_equals_single_engine)."""
    npx, npy, steps = 96, 150, 300
    single = run_slabs("ZIGZAG", solver, npx, npy, steps, 1, angle=20)[0]
    split = run_slabs("ZIGZAG", solver, npx, npy, steps, 3, angle=20)
    ext = (B.C.c_int32 * 4)()
    single.L.b200fdtd_get_lean_extent(single.engine.h, ext)
    assert tuple(ext) == (10, npx - 11, 10, npy - 11)
    split[1].L.b200fdtd_get_lean_extent(split[1].engine.h, ext)
    assert tuple(ext) == (10, npx - 11, split[1].j0 + 1, split[1].j0 + split[1].nj - 1)
    for slot in (0, 2, 3, 5, 6, 8):           # E, D, H, B (TM: Ez Dz Hx Bx Hy By; TE: Ex Dx Ey Dy Hz Bz)
        whole = single.gather_field(slot)
        parts = np.concatenate([r.gather_field(slot) for r in split], axis=1)
        assert np.abs(whole).max() > 0, slot
        assert rel_err(parts, whole) <= TOL_FIELD, slot
    for r in split + [single]:
        r.close()


def test_lean_switch_mid_run(plugin_lib, oracle):
    """Only differences of M / J enter the update outside the frame, so the form may be switched
    on and off between steps without leaving the tolerance."""
    n, steps, hu = 120, 450, 20
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, n, steps=steps, h_u_nm=hu)
    cpu = oracle_for(oracle, gpu, steps, hu=hu)
    h = gpu.engine_handle()
    for chunk, lean in ((150, 0), (150, 1), (150, 0)):
        gpu.sync()
        B.check(gpu.L.b200fdtd_set_option(h, B.OPT_LEAN_INTERIOR, lean), "set_option")
        gpu.step(chunk)
    cpu.step(steps)
    for f in ("Ez", "Hx", "Hy"):
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, f
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD


def test_lean_f32(plugin_lib, oracle):
    """Single precision + lean interior: the one-cell-per-thread kernels, float tolerance."""
    n, steps, hu = 120, 600, 20
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, steps=steps, h_u_nm=hu, precision="f32")
    assert gpu.lean_extent()[0] <= gpu.lean_extent()[1]
    cpu = oracle.OracleSim(oracle.TM, n, n, steps, gpu.eps(), h_u_nm=hu)
    gpu.run()
    cpu.step(steps)
    for f in ("Ez", "Hx", "Hy"):
        want = cpu.field(f)
        assert np.abs(gpu.field(f) - want).max() <= 2e-4 * np.abs(want).max(), f
    assert np.all(gpu.any_field(4)[30:80, 30:80] == 0)          # Mx untouched outside the frame
    assert rel_err(gpu.finish(), cpu.far_field()) <= 2e-3

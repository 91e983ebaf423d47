"""Parity of the CUDA path (through the C plugin surface and the engine C ABI) against
the CPU oracle, the golden vectors recorded from the unmodified reference, and -- when
oracle/_ref/libref.so travelled with the snapshot -- the reference itself.

Tolerances (max-abs-diff / max-abs-ref; BASELINE.json north_star):
  field snapshots  <= 1e-12      U/W arrays and far-field table  <= 1e-10
  permittivity / index maps: bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ANGLE_ROWS, LAMBDA_ROWS, TOL_FARFIELD, TOL_FIELD, bit_equal, golden, rel_err
from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun, split_columns

pytestmark = pytest.mark.gpu

MODEL_NAMES = {v: k for k, v in B.MODELS.items()}
TM_FIELDS = ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"]
TE_FIELDS = ["Ex", "Jx", "Dx", "Ey", "Jy", "Dy", "Hz", "Mz", "Bz"]


@pytest.fixture(autouse=True)
def _workdir(in_tmp_cwd, monkeypatch):
    monkeypatch.setenv("MPIFDTD_TRACE_IMAGE", os.path.join(os.path.dirname(__file__), "golden",
                                                           "traceImage_fixture.txt"))
    yield


def oracle_for(oracle, gpu, steps, hu=10, pml=10, lam=500, angle=0, point_source=False):
    """CPU oracle fed with the eps maps the plugin built (those are pinned bit-exactly
    against the reference in tests/test_materials_cpu.py)."""
    L = gpu.L
    npx, npy = gpu.n_px, gpu.n_py
    if gpu.solver == 2:
        return oracle.OracleSim(oracle.TM, npx, npy, steps, gpu.eps(), h_u_nm=hu, pml=pml, lambda_nm=lam,
                                angle_deg=angle, point_source=point_source)
    ex = gpu.eps()
    ey = np.empty((npx, npy))
    L.mpifdtd_fill_eps(ey.ctypes.data, 0.0, 0.5, B.D_X)
    return oracle.OracleSim(oracle.TE, npx, npy, steps, ex, ey, h_u_nm=hu, pml=pml, lambda_nm=lam,
                            angle_deg=angle, point_source=point_source)


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", ["mie_tm_upml_88x96", "mie_te_upml_88x96", "zigzag_tm_upml_72x120_a30",
                                  "layer_te_upml_80x110_a45"])
def test_gpu_matches_reference_golden_run(plugin_lib, name):
    g = golden(name + ".npz")
    npx, npy, hu, steps, angle, solver, model = (int(v) for v in g["meta"])
    gpu = B.Plugin(MODEL_NAMES[model], solver, npx, npy, steps=steps, h_u_nm=hu, angle_deg=angle)
    assert bit_equal(gpu.eps(), g["EPS_EZ" if solver == 2 else "EPS_EX"])
    half = steps // 2
    gpu.step(half)
    for key in [k for k in g.files if k.startswith("mid_")]:
        assert rel_err(gpu.field(key[4:]), g[key]) <= TOL_FIELD, key
    gpu.step(steps - half)
    for key in [k for k in g.files if k.startswith("end_")]:
        assert rel_err(gpu.field(key[4:]), g[key]) <= TOL_FIELD, key
    uw_order = ["Ux", "Uy", "Wz"] if solver == 2 else ["Wx", "Wy", "Uz"]
    first = True
    for key in [k for k in g.files if k.startswith("uw_")]:
        mine = gpu.ntff_uw(uw_order.index(key[3:]), project=first)[ANGLE_ROWS, :steps]
        first = False
        assert rel_err(mine, g[key]) <= TOL_FARFIELD, key
    table = gpu.finish()
    assert table.shape == (321, 360)
    assert rel_err(table[LAMBDA_ROWS, :], g["far_field_rows"]) <= TOL_FARFIELD
    assert os.path.getsize("%d[deg]_380nm_700nm_b.dat" % angle) == 321 * 360 * 8
    assert os.path.exists("%d[deg].txt" % angle)


# ---------------------------------------------------------------- all nine fields
@pytest.mark.parametrize("solver,names", [("TM_UPML_2D", TM_FIELDS), ("TE_UPML_2D", TE_FIELDS)])
def test_all_state_arrays_vs_oracle(plugin_lib, oracle, solver, names, monkeypatch):
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")       # keep all arraySize bins, spill included
    n, steps, hu = 120, 420, 20
    gpu = B.Plugin("MIE_CYLINDER", solver, n, n + 10, steps=steps, h_u_nm=hu)
    cpu = oracle_for(oracle, gpu, steps, hu=hu)
    gpu.run()
    cpu.step(steps)
    for slot, name in enumerate(names):
        assert rel_err(gpu.any_field(slot), cpu.field(slot)) <= TOL_FIELD, name
    for slot in range(3):
        mine = gpu.ntff_uw(slot, project=(slot == 0))
        assert mine.shape == (360, cpu.array_size)
        assert rel_err(mine, cpu.uw(slot)) <= TOL_FARFIELD, slot
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD


# ---------------------------------------------------------------- every model
@pytest.mark.parametrize("model,npx,npy", [("NO_MODEL", 64, 72), ("MIE_CYLINDER", 150, 150), ("LAYER", 100, 150),
                                           ("MORPHO_SCALE", 100, 190), ("ZIGZAG", 100, 160),
                                           ("TRACE_IMAGE", 110, 100)])
@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_each_material_model_vs_oracle(plugin_lib, oracle, model, npx, npy, solver):
    steps = 560
    gpu = B.Plugin(model, solver, npx, npy, steps=steps, angle_deg=10)
    cpu = oracle_for(oracle, gpu, steps, angle=10)
    gpu.run()
    cpu.step(steps)
    names = ["Ez", "Hx", "Hy"] if solver == "TM_UPML_2D" else ["Ex", "Ey", "Hz"]
    for f in names:
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, (model, f)
    if model == "NO_MODEL":      # every reference source is proportional to (eps0/eps - 1): all zero
        assert all(np.all(gpu.field(f) == 0) for f in names)
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD


def test_concentric_model_opt_in(plugin_lib, oracle, monkeypatch):
    monkeypatch.setenv("MPIFDTD_ENABLE_CONCENTRIC", "1")
    steps = 300
    gpu = B.Plugin("CONCENTRIC_CIRCLE", "TM_UPML_2D", 150, 150, steps=steps, h_u_nm=20)
    assert len(np.unique(gpu.eps())) > 3
    cpu = oracle_for(oracle, gpu, steps, hu=20)
    gpu.run()
    cpu.step(steps)
    assert rel_err(gpu.field("Ez"), cpu.field("Ez")) <= TOL_FIELD
    gpu.finish()


# ---------------------------------------------------------------- configs[0]: NoModel 256^2
def test_nomodel_256_point_source(plugin_lib, oracle):
    """BASELINE configs[0]: NoModel TM UPML 256 x 256.  The reference's scattered-field
    sources vanish for eps == eps0, so the opt-in point source (field_pointLight,
    field.c:145-152) drives the grid; the PML must absorb it."""
    steps = 400
    gpu = B.Plugin("NO_MODEL", "TM_UPML_2D", 256, steps=steps, point_source=True)
    cpu = oracle_for(oracle, gpu, steps, point_source=True)
    gpu.run()
    cpu.step(steps)
    ez = gpu.field("Ez")
    assert np.abs(ez).max() > 1e-3
    for f in ("Ez", "Hx", "Hy"):
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, f
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD
    B.lib().mpifdtd_enablePointSource(0)


# ---------------------------------------------------------------- configs[1] size
@pytest.mark.parametrize("solver,field", [("TM_UPML_2D", "Ez"), ("TE_UPML_2D", "Hz")])
def test_mie_1024_short_run_vs_oracle(plugin_lib, oracle, solver, field):
    steps = 40
    gpu = B.Plugin("MIE_CYLINDER", solver, 1024, steps=steps)
    cpu = oracle_for(oracle, gpu, steps)
    gpu.run()
    cpu.step(steps)
    assert rel_err(gpu.field(field), cpu.field(field)) <= TOL_FIELD
    for slot in range(3):
        assert rel_err(gpu.ntff_uw(slot, project=(slot == 0)), cpu.uw(slot)[:, :steps]) <= TOL_FARFIELD
    gpu.finish()


def test_one_pass_default_at_4096_vs_oracle(plugin_lib, oracle, monkeypatch, in_tmp_cwd):
    """A grid where the auto rule itself picks the one-pass step (form 3), with a material model and
    the NTFF box, compared DIRECTLY with the oracle -- all nine arrays and every bin of U/W -- instead
    of by transitivity through the two-kernel form."""
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")
    n, steps = 4096, 50
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, steps=steps)
    form = B.C.c_int32(-1)
    B.check(gpu.L.b200fdtd_get_step_form(gpu.engine_handle(), B.C.byref(form)), "get_step_form")
    assert form.value == 3
    cpu = oracle_for(oracle, gpu, steps)
    gpu.run()
    cpu.step(steps)
    assert np.abs(cpu.field(0)).max() > 0
    for slot in range(9):
        assert rel_err(gpu.any_field(slot), cpu.field(slot)) <= TOL_FIELD, slot
    # (50 steps do not carry the scattered field from the cylinder to the NTFF box 2000 cells away: the
    # surface history, and with it U/W, must be exactly as empty as the oracle's)
    for slot in range(3):
        assert rel_err(gpu.ntff_uw(slot, project=(slot == 0)), cpu.uw(slot)) <= TOL_FARFIELD, slot
    gpu.finish()


@pytest.mark.parametrize("solver,model", [("TM_UPML_2D", "MIE_CYLINDER"), ("TE_UPML_2D", "LAYER"),
                                          ("TM_UPML_2D", "MORPHO_SCALE")])
def test_palette_upload_equals_dense_upload(plugin_lib, monkeypatch, in_tmp_cwd, solver, model):
    """simulator_init ships the permittivity map as 16-bit indices + a table of its distinct values
    (b200fdtd_set_eps_palette) unless told otherwise: the device map, and with it every field, must
    be what the dense upload gives, bit for bit."""
    runs = {}
    for form in ("palette", "dense"):
        if form == "dense":
            monkeypatch.setenv("MPIFDTD_EPS_DENSE", "1")
        gpu = B.Plugin(model, solver, 150, 170, steps=300, h_u_nm=20, angle_deg=25)
        gpu.run()
        runs[form] = [gpu.any_field(s) for s in range(9)]
        gpu.finish()
    assert np.abs(runs["dense"][0]).max() > 0
    for a, b in zip(runs["palette"], runs["dense"]):
        assert bit_equal(a, b)


# ---------------------------------------------------------------- lifecycle
def test_reset_then_new_angle_matches_fresh_run(plugin_lib, oracle):
    """main.c:207-209: simulator_reset() writes the far field, zeroes state, then the
    driver changes the incidence angle and runs again with the same eps maps."""
    n, steps = 110, 500
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, steps=steps, h_u_nm=20)
    gpu.run()
    first = gpu.field("Ez")
    gpu.L.simulator_reset()
    assert os.path.exists("0[deg]_380nm_700nm_b.dat")
    assert np.all(gpu.field("Ez") == 0) and gpu.L.field_getTime() == 0.0
    gpu.L.field_setWaveAngle(25)
    gpu.info.angle_deg = 25
    cpu = oracle_for(oracle, gpu, steps, hu=20, angle=25)
    gpu.run()
    cpu.step(steps)
    assert rel_err(gpu.field("Ez"), cpu.field("Ez")) <= TOL_FIELD
    assert rel_err(gpu.field("Ez"), first) > 1e-3
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD
    assert os.path.exists("25[deg]_380nm_700nm_b.dat")


def test_init_after_finish_with_another_grid(plugin_lib, oracle):
    """main.c:198-205: finish, then init again with a different grid size."""
    for n, solver in [(80, "TE_UPML_2D"), (96, "TM_UPML_2D")]:
        gpu = B.Plugin("MIE_CYLINDER", solver, n, steps=50, h_u_nm=20)
        cpu = oracle_for(oracle, gpu, 50, hu=20)
        gpu.run()
        cpu.step(50)
        f = "Ez" if solver == "TM_UPML_2D" else "Hz"
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD
        gpu.finish()
        assert gpu.L.mpifdtd_upml_engine(gpu.solver) is None


def test_bitwise_determinism(plugin_lib):
    outs = []
    for _ in range(2):
        gpu = B.Plugin("ZIGZAG", "TM_UPML_2D", 100, 160, steps=300)
        gpu.run()
        outs.append((gpu.field("Ez"), gpu.ntff_uw(2), gpu.finish()))
    assert bit_equal(outs[0][0].view(np.float64), outs[1][0].view(np.float64))
    assert bit_equal(outs[0][1].view(np.float64), outs[1][1].view(np.float64))
    assert bit_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("npx,npy,steps,pml,lam,hu", [(40, 41, 1, 10, 500, 10), (64, 200, 90, 15, 633, 10),
                                                      (150, 157, 90, 12, 450, 5), (33, 37, 64, 5, 500, 20)])
def test_edge_shapes(plugin_lib, oracle, npx, npy, steps, pml, lam, hu):
    """Smallest workable box, one step, non-square grids, other pml / lambda / cell size.
    The point source makes sure something non-trivial propagates on tiny grids.
    (Grids wider than tall are not a valid reference configuration: RFperC comes from
    the y extent only, field.c:140-141, so its time shifts go far negative and the
    reference writes out of bounds -- see test_wide_grid_fields_only.)"""
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", npx, npy, steps=steps, h_u_nm=hu, pml=pml, lambda_nm=lam,
                   point_source=True)
    cpu = oracle_for(oracle, gpu, steps, hu=hu, pml=pml, lam=lam, point_source=True)
    gpu.run()
    cpu.step(steps)
    for f in ("Ez", "Hx", "Hy"):
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, f
    assert rel_err(gpu.finish(), cpu.far_field()) <= TOL_FARFIELD
    B.lib().mpifdtd_enablePointSource(0)


def test_wide_grid_fields_only(plugin_lib, oracle):
    """N_PX > N_PY: the reference's NTFF indexing is undefined there, the stencil is not.
    Fields must still match; the GPU projection simply drops bins below zero."""
    npx, npy, steps = 200, 64, 120
    gpu = B.Plugin("LAYER", "TM_UPML_2D", npx, npy, steps=steps, point_source=True)
    cpu = oracle_for(oracle, gpu, steps, point_source=True)
    gpu.run()
    cpu.step(steps, with_ntff=False)
    for f in ("Ez", "Hx", "Hy"):
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, f
    assert np.all(np.isfinite(gpu.finish()))
    B.lib().mpifdtd_enablePointSource(0)


def test_engine_call_order_is_enforced(plugin_lib):
    eng = B.Engine(2, 64, 64, 10)
    args = B.StepArgs()
    with pytest.raises(B.EngineError, match="before set_upml_tables"):
        eng.step(args)
    eng.close()


# ---------------------------------------------------------------- y-slab split on one GPU
def run_slabs(model, solver, npx, npy, steps, world, angle=0, precision=0):
    """`world` engines on one GPU exchanging halo columns through device buffers: the
    multi-GPU data path minus NCCL."""
    import torch
    L = B.lib()
    runs = [SlabRun(model, solver, npx, npy, steps, rank=r, world=world, device=0, angle_deg=angle,
                    precision=precision)
            for r in range(world)]
    bufs = [torch.zeros(2 * npx, dtype=torch.float64, device="cuda") for _ in range(world)]
    args = B.StepArgs()
    kind = runs[0].kind
    for _ in range(steps):
        L.mpifdtd_upml_step_args(kind, 0, C.byref(args))
        for which, phase in ((0, "phase_h"), (1, "phase_e")):
            for r in runs:
                getattr(r.engine, phase)(args)
            for r in runs:
                r.engine.halo_pack(which, bufs[r.rank].data_ptr())
            for r in runs:
                r.engine.sync()
            for r in runs:
                src = r.rank - 1 if which == 0 else r.rank + 1
                if 0 <= src < world:
                    r.engine.halo_unpack(which, bufs[src].data_ptr())
            for r in runs:
                r.engine.sync()
        for r in runs:
            r.engine.phase_sample(args)
        L.field_nextStep()
    return runs


@pytest.mark.parametrize("solver,world", [("TM_UPML_2D", 2), ("TM_UPML_2D", 3), ("TE_UPML_2D", 2)])
def test_slab_split_equals_single_engine(plugin_lib, solver, world):
    npx, npy, steps = 96, 150, 330
    single = run_slabs("ZIGZAG", solver, npx, npy, steps, 1, angle=20)[0]
    split = run_slabs("ZIGZAG", solver, npx, npy, steps, world, angle=20)
    for slot in range(9):
        whole = single.gather_field(slot)
        parts = np.concatenate([r.gather_field(slot) for r in split], axis=1)
        assert bit_equal(parts.view(np.float64), whole.view(np.float64)), slot
    single.project()
    want = [single.engine.uw(s) for s in range(3)]
    for r in split:
        r.project()
    for s in range(3):
        total = sum(r.engine.uw(s) for r in split)
        assert rel_err(total, want[s]) <= 1e-13, s
    for r in split + [single]:
        r.close()


# ---------------------------------------------------------------- full-size property
def test_linearity_at_full_size_16384(plugin_lib):
    """BASELINE configs[4] size (16384 x 16384 on one GPU): the update is linear, so
    doubling the initial state doubles the result exactly (scaling by 2 is exact in
    binary floating point).  Checked on the whole Ez plane after 3 steps."""
    import torch
    if torch.cuda.mem_get_info()[1] < 100e9:
        pytest.skip("needs a >= 100 GB GPU")
    n, steps = 16384, 3
    rng = np.random.default_rng(11)
    seed_rows = (rng.standard_normal((64, n)) + 1j * rng.standard_normal((64, n)))
    results = []
    for scale in (1.0, 2.0):
        run = SlabRun("NO_MODEL", "TM_UPML_2D", n, n, steps, with_ntff=False)
        ez0 = np.zeros((n, n), dtype=np.complex128)
        ez0[n // 2 - 32:n // 2 + 32, :] = scale * seed_rows
        ez0[:, 0] = ez0[:, -1] = 0
        run.engine.set_field(0, ez0)
        del ez0
        for _ in range(steps):
            run.step()
        results.append(run.gather_field(0))
        run.close()
    assert np.abs(results[0]).max() > 0.1
    assert np.array_equal(results[1], 2.0 * results[0])
    # nothing can have travelled more than `steps` cells from the seeded band
    assert np.all(results[0][: n // 2 - 32 - steps - 1, :] == 0)


# ---------------------------------------------------------------- live reference
def test_far_field_vs_live_reference(plugin_lib):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    n, steps = 200, 600
    cwd = os.getcwd()
    ref = reflib.RefSim("MIE_CYLINDER", "TE_UPML_2D", n, steps=steps)
    ref.run()
    ref_hz = ref.field("Hz")
    ref_eps = ref.coef("EPS_EX")
    want = ref.finish()
    os.chdir(cwd)
    gpu = B.Plugin("MIE_CYLINDER", "TE_UPML_2D", n, steps=steps)
    assert bit_equal(gpu.eps(), ref_eps)
    gpu.run()
    assert rel_err(gpu.field("Hz"), ref_hz) <= TOL_FIELD
    assert rel_err(gpu.finish(), want) <= TOL_FARFIELD

"""Optional single-precision path of the UPML solvers (north_star: "an optional float path with
its own tolerance"): complex64 fields, f32 permittivity and coefficients on the GPU; sources,
NTFF history and far-field post-processing stay double.

Tolerances of this path (max-abs-diff / max-abs-ref against the double CPU oracle):
  field snapshots <= 2e-4, far-field table <= 2e-3 after a few hundred steps
(float rounding, ~6e-8 per operation, accumulated in the J/D and M/B recurrences)."""
import os

import numpy as np
import pytest

from helpers import rel_err
from mpifdtd_b200 import binding as B
from mpifdtd_b200.slab import SlabRun

pytestmark = pytest.mark.gpu
TOL_FIELD_F32 = 2e-4
TOL_FARFIELD_F32 = 2e-3


@pytest.fixture(autouse=True)
def _restore_precision(in_tmp_cwd):
    yield
    B.lib().mpifdtd_setPrecision(0)


@pytest.mark.parametrize("solver,model,angle", [("TM_UPML_2D", "MIE_CYLINDER", 0), ("TE_UPML_2D", "MIE_CYLINDER", 0),
                                                ("TM_UPML_2D", "ZIGZAG", 30), ("TE_UPML_2D", "LAYER", 45)])
def test_f32_fields_and_far_field_vs_oracle(plugin_lib, oracle, solver, model, angle):
    n, steps, hu = 120, 600, 20
    gpu = B.Plugin(model, solver, n, steps=steps, h_u_nm=hu, angle_deg=angle, precision="f32")
    L = gpu.L
    if gpu.solver == 2:
        cpu = oracle.OracleSim(oracle.TM, n, n, steps, gpu.eps(), h_u_nm=hu, angle_deg=angle)
        names = ("Ez", "Hx", "Hy")
    else:
        ey = np.empty((n, n))
        L.mpifdtd_fill_eps(ey.ctypes.data, 0.0, 0.5, B.D_X)
        cpu = oracle.OracleSim(oracle.TE, n, n, steps, gpu.eps(), ey, h_u_nm=hu, angle_deg=angle)
        names = ("Ex", "Ey", "Hz")
    gpu.run()
    cpu.step(steps)
    scale = max(np.abs(cpu.field(f)).max() for f in names[:2] if f[0] == "E")
    assert scale > 1e-3
    for f in names:
        want = cpu.field(f)
        ref_scale = np.abs(want).max()
        assert np.abs(gpu.field(f) - want).max() <= TOL_FIELD_F32 * ref_scale, f
    # and it really is the float path: a double run agrees with the oracle ~1e9 times better
    assert rel_err(gpu.field(names[0]), cpu.field(names[0])) > 1e-9
    far = gpu.finish()
    assert rel_err(far, cpu.far_field()) <= TOL_FARFIELD_F32


def test_f32_auxiliary_arrays_and_setters_roundtrip(plugin_lib):
    """get_field/set_field widen and narrow through the double staging plane."""
    n = 64
    run = SlabRun("MIE_CYLINDER", "TM_UPML_2D", n, n, 8, with_ntff=False, precision="f32")
    rng = np.random.default_rng(5)
    a = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64).astype(np.complex128)
    run.engine.set_field(4, a)
    assert np.array_equal(run.engine.get_field(4), a)
    assert run.engine.device_bytes() < 0.6 * 16 * 9 * (n + 2) * (n + 16) + 4e5      # float2 planes, not double2
    run.close()


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_f32_slab_split_equals_single_engine(plugin_lib, solver):
    """y-slab halos (pack/unpack path) in single precision: bit-identical to one slab."""
    from test_gpu_parity import run_slabs
    npx, npy, steps = 72, 96, 120
    single = run_slabs("MIE_CYLINDER", solver, npx, npy, steps, 1, angle=20, precision="f32")[0]
    split = run_slabs("MIE_CYLINDER", solver, npx, npy, steps, 3, angle=20, precision="f32")
    for slot in range(9):
        whole = single.gather_field(slot)
        parts = np.concatenate([r.gather_field(slot) for r in split], axis=1)
        assert np.array_equal(parts.view(np.float64), whole.view(np.float64)), slot
    assert np.abs(single.gather_field(0)).max() > 0
    for r in split + [single]:
        r.close()


def test_f32_rejected_where_not_built(plugin_lib):
    import ctypes as C
    grid = B.Grid(0, 64, 64, 10, 0, 64, 1, 62, 1, 62, -1, 1, B.MU_0_S)      # split-field kind in f32
    h = C.c_void_p()
    assert plugin_lib.b200fdtd_create(C.byref(grid), C.byref(h)) == 1
    grid = B.Grid(2, 64, 64, 10, 0, 64, 1, 62, 1, 62, -1, 7, B.MU_0_S)      # unknown precision
    assert plugin_lib.b200fdtd_create(C.byref(grid), C.byref(h)) == 1


@pytest.mark.parametrize("solver,model,npx,npy,batch", [
    ("TM_UPML_2D", "MIE_CYLINDER", 120, 120, None), ("TE_UPML_2D", "MIE_CYLINDER", 120, 121, None),
    ("TM_UPML_2D", "ZIGZAG", 97, 301, None), ("TE_UPML_2D", "LAYER", 301, 98, None),
    ("TM_UPML_2D", "MIE_CYLINDER", 110, 113, [0, 30, 65])])
def test_f32_pair_kernels_are_bit_identical_to_one_cell_kernels(plugin_lib, monkeypatch, solver, model, npx, npy, batch):
    """Two cells per thread with 128-bit accesses (upml_pairs_f32.cuh) share the arithmetic of the
    one-cell kernels: every array must agree bit for bit, odd and even widths alike."""
    steps = 540
    results = {}
    for pairs in ("0", "1"):
        monkeypatch.setenv("B200FDTD_F32_PAIRS", pairs)
        monkeypatch.setenv("MPIFDTD_DEFER_STEPS", "0")
        gpu = B.Plugin(model, solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=25, precision="f32", angle_batch=batch)
        gpu.run()
        if batch:
            gpu.select_angle(2)
        results[pairs] = [gpu.any_field(s) for s in range(9)] + [gpu.ntff_uw(s, project=(s == 0)) for s in range(3)]
        gpu.finish()
    assert np.abs(results["0"][0]).max() > 1e-4
    for n, (a, b) in enumerate(zip(results["0"], results["1"])):
        assert np.array_equal(a.view(np.float64), b.view(np.float64)), n

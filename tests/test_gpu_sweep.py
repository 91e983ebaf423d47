"""Angle sweeps as one batched GPU job (replaces the one-angle-per-rank loop
of main.c:114-138,183-211).  A batched engine runs the same kernels with the same arithmetic
on every simulation, so each angle of a batch must be BIT-IDENTICAL to the unbatched run at
that angle -- fields and far-field files alike -- which in turn is pinned to the reference by
tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from helpers import TOL_FARFIELD, TOL_FIELD, bit_equal, rel_err
from mpifdtd_b200 import binding as B

pytestmark = pytest.mark.gpu
FIELDS = {2: ("Ez", "Hx", "Hy", "Jz", "Bx"), 3: ("Ex", "Ey", "Hz", "Dy", "Mz")}


@pytest.fixture(autouse=True)
def _unbatched_afterwards(in_tmp_cwd):
    yield
    B.lib().mpifdtd_setAngleBatch(None, 0)
    B.lib().mpifdtd_setPrecision(0)


@pytest.mark.parametrize("solver,model,precision", [(2, "MIE_CYLINDER", "f64"), (3, "MIE_CYLINDER", "f64"),
                                                    (2, "ZIGZAG", "f64"), (3, "LAYER", "f32")])
def test_batched_angles_equal_single_runs(plugin_lib, solver, model, precision):
    npx, npy, steps, angles = 104, 120, 520, [0, 25, 90, 135]
    singles = {}
    for ang in angles:
        gpu = B.Plugin(model, solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=ang, precision=precision)
        gpu.run()
        fields = {f: gpu.field(f) for f in FIELDS[solver]}
        singles[ang] = (fields, gpu.finish())
    assert np.abs(singles[25][0][FIELDS[solver][0]]).max() > 1e-3
    assert not np.array_equal(singles[0][1], singles[90][1])          # the angles really differ

    batch = B.Plugin(model, solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=angles[0], precision=precision,
                     angle_batch=angles)
    launches0 = batch.launches()
    batch.run()
    assert batch.launches() - launches0 <= 4 * steps                  # still H, E, sample (+ clock) per step
    for k, ang in enumerate(angles):
        batch.select_angle(k)
        for f in FIELDS[solver]:
            assert bit_equal(batch.field(f), singles[ang][0][f]), (ang, f)
    batch.finish()
    for ang in angles:
        assert bit_equal(batch.far_field_file(ang), singles[ang][1]), ang
        assert os.path.exists("%d[deg].txt" % ang)


SPLIT_NAMES = {0: ("Ez", "Hx", "Hy", "Ezx", "Ezy"), 1: ("Ex", "Ey", "Hz", "Hzx", "Hzy")}
SPLIT_NAMES[6], SPLIT_NAMES[7] = SPLIT_NAMES[0], SPLIT_NAMES[1]
SPLIT_DUMP = {0: "tm_20nm.txt", 1: "te_20nm.txt", 6: "ns_tm_20nm.txt", 7: "ns_te_20nm.txt"}


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("solver", [0, 1, 6, 7])
def test_batched_angles_of_the_split_field_solvers(plugin_lib, monkeypatch, solver, dense):
    """main.c as shipped sweeps the incidence angle with NS_TE_2D (main.c:152,183-211); ids 0, 1, 6, 7 as
    one batched engine: every simulation bit-identical to its own single run (CW source with the
    per-angle wave vector and, for id 7, the polarisation factors and on/off switches of
    nsFdtdTE.c:243-250), the validation-circle dump of every angle on disk."""
    if dense:
        monkeypatch.setenv("MPIFDTD_SPLIT_DENSE", "1")
    npx, npy, steps, angles = 150, 340, 260, [0, 30, 90, 200]
    singles = {}
    for ang in angles:
        gpu = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=ang)
        gpu.run()
        singles[ang] = {f: gpu.field(f) for f in SPLIT_NAMES[solver]}
        gpu.finish()
        singles[ang]["dump"] = open(SPLIT_DUMP[solver]).read()
    assert np.abs(singles[30][SPLIT_NAMES[solver][0]]).max() > 1e-6
    assert not np.array_equal(singles[0][SPLIT_NAMES[solver][0]], singles[90][SPLIT_NAMES[solver][0]])

    batch = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=angles[0], angle_batch=angles)
    launches0 = batch.launches()
    batch.run()
    assert batch.launches() - launches0 == 2 * steps                  # still two kernels per step
    for k, ang in enumerate(angles):
        batch.select_angle(k)
        for f in SPLIT_NAMES[solver]:
            assert bit_equal(batch.field(f), singles[ang][f]), (ang, f)
    batch.finish()
    for ang in angles:
        assert open("%d[deg]_%s" % (ang, SPLIT_DUMP[solver])).read() == singles[ang]["dump"], ang
    assert open(SPLIT_DUMP[solver]).read() == singles[angles[-1]]["dump"]       # upstream's name: the last angle


def test_angle_sweep_of_the_shipped_default_solver(plugin_lib):
    """mpifdtd_runAngleSweep with NS_TE_2D: chunks of 3 and the one-at-a-time loop leave the same dump."""
    L = B.lib()
    n, steps = 120, 150
    L.models_setModel(B.MODELS["MIE_CYLINDER"])
    L.simulator_setSolver(7)
    info = B.FieldInfo(n * 20, n * 20, 20, 10, 500, 0, steps)
    assert L.mpifdtd_runAngleSweep(info, 0, 80, 20, 3) == 5
    swept = {a: open("%d[deg]_ns_te_20nm.txt" % a).read() for a in range(0, 81, 20)}
    last = open("ns_te_20nm.txt").read()
    assert last == swept[80] and swept[0] != swept[80]
    os.remove("ns_te_20nm.txt")
    assert L.mpifdtd_runAngleSweep(info, 0, 80, 20, 1) == 5
    assert open("ns_te_20nm.txt").read() == last


def test_batch_vs_oracle_directly(plugin_lib, oracle):
    """One angle of a batch against the CPU oracle itself (not only against our own single run)."""
    n, steps, angles = 96, 400, [10, 60]
    batch = B.Plugin("MIE_CYLINDER", 2, n, steps=steps, h_u_nm=20, angle_batch=angles)
    eps = batch.eps()
    batch.run()
    batch.select_angle(1)
    cpu = oracle.OracleSim(oracle.TM, n, n, steps, eps, h_u_nm=20, angle_deg=60)
    cpu.step(steps)
    for f in ("Ez", "Hx", "Hy"):
        assert rel_err(batch.field(f), cpu.field(f)) <= TOL_FIELD, f
    batch.finish()
    assert rel_err(batch.far_field_file(60), cpu.far_field()) <= TOL_FARFIELD


def test_reset_of_a_batch_restarts_every_angle(plugin_lib):
    n, steps, angles = 88, 200, [0, 45, 75]
    batch = B.Plugin("MIE_CYLINDER", 2, n, steps=steps, h_u_nm=20, angle_batch=angles)
    batch.run()
    first = []
    for k in range(len(angles)):
        batch.select_angle(k)
        first.append(batch.field("Ez"))
    L = batch.L
    L.simulator_reset()                    # writes every angle's files, zeroes every simulation
    L.field_reset()
    tables = [batch.far_field_file(a) for a in angles]
    batch.select_angle(2)
    assert np.all(batch.field("Ez") == 0)
    batch.run()
    for k in range(len(angles)):
        batch.select_angle(k)
        assert bit_equal(batch.field("Ez"), first[k]), k
    batch.finish()
    for a, t in zip(angles, tables):
        assert bit_equal(batch.far_field_file(a), t), a


def test_run_angle_sweep_chunks_and_names(plugin_lib):
    """mpifdtd_runAngleSweep: 0..60 step 15 in chunks of at most 2 -> 3 batches, 5 simulations,
    every angle's files present and equal to a plain single run."""
    L = B.lib()
    n, steps = 80, 160
    L.models_setModel(B.MODELS["MIE_CYLINDER"])
    L.simulator_setSolver(2)
    info = B.FieldInfo(n * 20, n * 20, 20, 10, 500, 0, steps)
    assert L.mpifdtd_runAngleSweep(info, 0, 60, 15, 2) == 5
    first = {a: np.fromfile("%d[deg]_380nm_700nm_b.dat" % a).reshape(321, 360) for a in range(0, 61, 15)}
    assert L.mpifdtd_runAngleSweep(info, 0, 60, 15, 1) == 5          # the reference's own one-at-a-time loop
    for a in first:
        assert bit_equal(np.fromfile("%d[deg]_380nm_700nm_b.dat" % a).reshape(321, 360), first[a]), a
    swept = {a: np.fromfile("%d[deg]_380nm_700nm_b.dat" % a).reshape(321, 360) for a in range(0, 61, 15)}
    os.makedirs("single", exist_ok=True)
    os.chdir("single")
    gpu = B.Plugin("MIE_CYLINDER", 2, n, steps=steps, h_u_nm=20, angle_deg=45)
    gpu.run()
    assert bit_equal(gpu.finish(), swept[45])


def test_batch_rejected_where_not_built(plugin_lib):
    import ctypes as C
    h = C.c_void_p()
    for kind, j0, nj in ((4, 0, 64), (2, 0, 32), (7, 0, 32)):          # MPI variant, slabs
        grid = B.Grid(kind, 64, 64, 10, j0, nj, 1, 62, 1, 62, -1, 0, B.MU_0_S, 4, 0)
        assert plugin_lib.b200fdtd_create(C.byref(grid), C.byref(h)) == 1, kind
    grid = B.Grid(2, 64, 64, 10, 0, 64, 1, 62, 1, 62, -1, 0, B.MU_0_S, 3, 0)
    assert plugin_lib.b200fdtd_create(C.byref(grid), C.byref(h)) == 0
    args = B.StepArgs()
    assert plugin_lib.b200fdtd_step(h, C.byref(args)) == 4               # ERR_STATE: nothing uploaded yet
    assert plugin_lib.b200fdtd_select_batch(h, 3) == 1
    plugin_lib.b200fdtd_destroy(h)


def test_sweep_tool_directory_chain_and_rank_striding(plugin_lib, tmp_path):
    """mpifdtd_b200/mpifdtd_sweep: main.c's batch mode (structure loop, moveDir chain) with each
    structure's angles as one batched engine; two 'ranks' stride through the angle list like
    the reference's MPI ranks (main.c:126-138)."""
    import subprocess
    tool = os.path.join(os.path.dirname(B.LIB_PATH), "mpifdtd_sweep")
    assert os.path.exists(tool), "run __graft_entry__.build()"
    cfg = tmp_path / "config.txt"
    # width height h_u pml lambda steps startAngle endAngle deltaAngle model solver
    cfg.write_text("# sweep test\n2000\n2000\n20\n10\n500\n160\n0\n50\n10\n1\n2\n")
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2")
        p = subprocess.run([tool, str(cfg)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:]
        assert "rank %d of 2 ran 3 simulation(s) over 1 structure(s)" % rank in p.stdout
    out = tmp_path / "MieCylinderModel" / "radius_500nm" / "hu_20nm" / "TM_UPML"
    names = sorted(os.listdir(out))
    assert [n for n in names if n.endswith("_b.dat")] == sorted("%d[deg]_380nm_700nm_b.dat" % a for a in range(0, 51, 10))
    os.chdir(tmp_path)
    gpu = B.Plugin("MIE_CYLINDER", 2, 100, steps=160, h_u_nm=20, angle_deg=30)
    gpu.run()
    want = gpu.finish()
    got = np.fromfile(str(out / "30[deg]_380nm_700nm_b.dat")).reshape(321, 360)
    assert bit_equal(got, want)

"""Multi-step replay: b200fdtd_run_steps captures [H phase, E phase + source, surface sample,
clock] for a chunk of steps into a CUDA graph, the step's time coming from a device-side clock,
and the plugin's update() hands its steps over in such chunks (deferred stepping).  Everything
must be bit-identical to stepping one launch sequence at a time."""
import numpy as np
import pytest

from helpers import bit_equal
from mpifdtd_b200 import binding as B

pytestmark = pytest.mark.gpu
FIELDS = {2: ("Ez", "Hx", "Hy", "Jz", "Dz", "Mx", "Bx", "My", "By"), 3: ("Ex", "Ey", "Hz", "Jx", "Dx", "Jy", "Dy", "Mz", "Bz")}


@pytest.fixture(autouse=True)
def _clean(in_tmp_cwd, monkeypatch):
    yield
    B.lib().mpifdtd_setAngleBatch(None, 0)
    B.lib().mpifdtd_setPrecision(0)


def run(solver, model, n, steps, angle, precision="f64", batch=None, mid=None):
    gpu = B.Plugin(model, solver, n, steps=steps, h_u_nm=20, angle_deg=angle, precision=precision,
                   angle_batch=batch)
    snaps = {}
    if mid:
        gpu.step(mid)
        snaps["mid"] = gpu.field(FIELDS[gpu.solver][0])          # a getter in the middle of a run flushes
        gpu.step(steps - mid)
    else:
        gpu.run()
    if batch:
        gpu.select_angle(len(batch) - 1)
    snaps.update({f: gpu.field(f) for f in FIELDS[gpu.solver]})
    uw = [gpu.ntff_uw(s, project=(s == 0)) for s in range(3)]
    n_launch = gpu.launches()
    far = gpu.finish()
    return snaps, uw, far, n_launch


@pytest.mark.parametrize("solver,model,precision,chunk,batch", [
    ("TM_UPML_2D", "MIE_CYLINDER", "f64", None, None), ("TE_UPML_2D", "MIE_CYLINDER", "f64", None, None),
    ("TM_UPML_2D", "ZIGZAG", "f64", "50", None),           # a chunk that does not divide the run
    ("TE_UPML_2D", "LAYER", "f32", "7", None),             # below the graph threshold: plain launches off the device clock
    ("TM_UPML_2D", "MIE_CYLINDER", "f64", "128", [0, 40, 75])])
def test_deferred_stepping_is_bit_identical(plugin_lib, monkeypatch, solver, model, precision, chunk, batch):
    n, steps, angle = 110, 530, 25
    monkeypatch.setenv("MPIFDTD_DEFER_STEPS", "0")
    want, want_uw, want_far, n_plain = run(solver, model, n, steps, angle, precision, batch, mid=200)
    monkeypatch.delenv("MPIFDTD_DEFER_STEPS")
    if chunk:
        monkeypatch.setenv("MPIFDTD_DEFER_CHUNK", chunk)
    got, got_uw, got_far, n_replay = run(solver, model, n, steps, angle, precision, batch, mid=200)
    assert np.abs(want[FIELDS[B.SOLVERS[solver]][0]]).max() > 1e-4
    for key in want:
        assert bit_equal(got[key], want[key]), key
    for s in range(3):
        assert bit_equal(got_uw[s], want_uw[s]), s
    assert bit_equal(got_far, want_far)
    assert n_replay >= n_plain                       # same kernels + one clock tick per step


SPLIT_FIELDS = {0: ("Ez", "Hx", "Hy", "Ezx", "Ezy"), 1: ("Ex", "Ey", "Hz", "Hzx", "Hzy")}
SPLIT_FIELDS[6], SPLIT_FIELDS[7] = SPLIT_FIELDS[0], SPLIT_FIELDS[1]


@pytest.mark.parametrize("solver,chunk,batch", [(0, None, None), (1, "100", None), (6, "50", None), (7, None, None),
                                               (7, "5", None), (7, "128", [0, 40, 75]), (0, "64", [10, 100])])
def test_deferred_stepping_of_the_split_field_solvers(plugin_lib, monkeypatch, solver, chunk, batch):
    """b200fdtd_run_split_steps: update() records the step's CW arguments, chunks of them replay from
    one CUDA graph whose kernels read their own step's record.  Bit-identical to one b200fdtd_step
    per update(), a getter in mid-run included; same number of kernels."""
    npx, npy, steps, mid = 130, 300, 430, 170

    def once():
        gpu = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=35, angle_batch=batch)
        gpu.step(mid)
        snaps = {"mid": gpu.field(SPLIT_FIELDS[solver][0])}
        gpu.step(steps - mid)
        if batch:
            gpu.select_angle(len(batch) - 1)
        snaps.update({f: gpu.field(f) for f in SPLIT_FIELDS[solver]})
        n_launch = gpu.launches()
        gpu.finish()
        return snaps, n_launch

    monkeypatch.setenv("MPIFDTD_DEFER_STEPS", "0")
    want, n_plain = once()
    monkeypatch.delenv("MPIFDTD_DEFER_STEPS")
    if chunk:
        monkeypatch.setenv("MPIFDTD_DEFER_CHUNK", chunk)
    got, n_replay = once()
    assert np.abs(want["mid"]).max() > 1e-6
    for key in want:
        assert bit_equal(got[key], want[key]), key
    assert n_replay == n_plain and 0 <= n_plain - 2 * steps <= 2      # + the fill kernels of init()


def test_run_steps_argument_checks(plugin_lib):
    import ctypes as C
    from mpifdtd_b200.slab import SlabRun
    r = SlabRun("MIE_CYLINDER", "TM_UPML_2D", 64, 64, 20)
    L, h = r.L, r.engine.h
    assert L.b200fdtd_run_steps(h, 0.0, 5) == 4                   # ERR_STATE: no pulse records uploaded
    src = (C.c_char * L.b200fdtd_struct_size(5))()
    assert L.b200fdtd_set_batch_sources(h, src) == 0
    assert L.b200fdtd_run_steps(h, 0.0, 5) == 0
    assert L.b200fdtd_run_steps(h, 18.0, 5) == 1                  # would run past the NTFF history
    assert L.b200fdtd_run_steps(h, 0.0, -1) == 1
    r.close()
    split = B.Plugin("MIE_CYLINDER", 0, 200, steps=4)
    assert L.b200fdtd_run_steps(split.engine_handle(), 0.0, 2) == 1   # split-field kinds: b200fdtd_run_split_steps
    L.b200fdtd_run_split_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    assert L.b200fdtd_run_split_steps(split.engine_handle(), None, 2) == 1
    args = (B.StepArgs * 3)()
    assert L.b200fdtd_run_split_steps(split.engine_handle(), args, -1) == 1
    assert L.b200fdtd_run_split_steps(split.engine_handle(), args, 3) == 0   # sources disabled: three plain steps
    split.finish()

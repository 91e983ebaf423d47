"""Unit-coefficient interior kernels (B200FDTD_OPT_UNIT_SPLIT, the default form on large grids).

Inside the frame-free rectangle every UPML coefficient is exactly 1.0 and 1.0 * x == x, so the
dedicated interior kernels evaluate the reference's expressions without table reads and
multiplications.  The claim is BIT-IDENTITY with the one-kernel-per-phase form, on every state
array, which is what these tests check (forcing the split on at test sizes, where `auto` would
not choose it)."""
import numpy as np
import pytest

from helpers import TOL_FIELD, bit_equal, rel_err
from mpifdtd_b200 import binding as B
from test_gpu_parity import oracle_for, run_slabs

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cwd(in_tmp_cwd):
    yield
    B.lib().mpifdtd_setPrecision(0)


def run_plugin(monkeypatch, split, model, solver, npx, npy, steps, angle, **kw):
    monkeypatch.setenv("B200FDTD_UNIT_SPLIT", str(split))
    gpu = B.Plugin(model, solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=angle, **kw)
    gpu.run()
    state = [gpu.any_field(s) for s in range(9)]
    far = gpu.finish()
    return state, far, gpu


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D", "MPI_TM_UPML_2D", "MPI_TE_UPML_2D"])
@pytest.mark.parametrize("model,angle", [("MIE_CYLINDER", 0), ("ZIGZAG", 30)])
def test_unit_split_is_bit_identical(plugin_lib, monkeypatch, solver, model, angle):
    npx, npy, steps = 150, 170, 500
    one, far_one, _ = run_plugin(monkeypatch, 0, model, solver, npx, npy, steps, angle)
    two, far_two, _ = run_plugin(monkeypatch, 1, model, solver, npx, npy, steps, angle)
    assert np.abs(one[0]).max() > 0
    for slot in range(9):
        assert bit_equal(two[slot], one[slot]), slot
    if far_one is not None:
        assert bit_equal(far_two, far_one)


def test_unit_split_vs_oracle_and_launch_count(plugin_lib, oracle, monkeypatch):
    """And against the CPU oracle directly; the split really launches the extra kernels."""
    monkeypatch.setenv("B200FDTD_UNIT_SPLIT", "1")
    monkeypatch.setenv("MPIFDTD_DEFER_STEPS", "0")
    n, steps = 120, 300
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, n, steps=steps, h_u_nm=20)
    cpu = oracle_for(oracle, gpu, steps, hu=20)
    before = gpu.launches()
    gpu.step(10)
    assert gpu.launches() - before == 10 * 5          # interior + frame per phase, NTFF sample
    gpu.step(steps - 10)
    cpu.step(steps)
    for slot in range(9):
        assert rel_err(gpu.any_field(slot), cpu.field(slot)) <= TOL_FIELD, slot
    gpu.finish()


def test_unit_split_store_h_and_batch(plugin_lib, monkeypatch):
    """STORE_H form (H arrays written every step) and an angle batch through the split."""
    res = {}
    for split in (0, 1):
        monkeypatch.setenv("B200FDTD_UNIT_SPLIT", str(split))
        monkeypatch.setenv("B200FDTD_STORE_H", "1")
        gpu = B.Plugin("MIE_CYLINDER", "TE_UPML_2D", 130, 140, steps=300, h_u_nm=20, angle_batch=[0, 40, 90])
        gpu.run()
        out = []
        for k in range(3):
            gpu.select_angle(k)
            out.append([gpu.any_field(s) for s in range(9)])
        gpu.finish()
        res[split] = out
    for k in range(3):
        for slot in range(9):
            assert bit_equal(res[1][k][slot], res[0][k][slot]), (k, slot)


@pytest.mark.parametrize("solver", ["TM_UPML_2D", "TE_UPML_2D"])
def test_unit_split_slabs_equal_single_engine(plugin_lib, solver, monkeypatch):
    """Bit-identical arithmetic => a y-slab split still reproduces the single engine bit for bit."""
    monkeypatch.setenv("B200FDTD_UNIT_SPLIT", "1")
    npx, npy, steps = 96, 150, 250
    single = run_slabs("ZIGZAG", solver, npx, npy, steps, 1, angle=20)[0]
    split = run_slabs("ZIGZAG", solver, npx, npy, steps, 3, angle=20)
    monkeypatch.setenv("B200FDTD_UNIT_SPLIT", "0")
    plain = run_slabs("ZIGZAG", solver, npx, npy, steps, 1, angle=20)[0]
    for slot in range(9):
        whole = single.gather_field(slot)
        parts = np.concatenate([r.gather_field(slot) for r in split], axis=1)
        assert bit_equal(parts, whole), slot
        assert bit_equal(plain.gather_field(slot), whole), slot
    assert np.abs(single.gather_field(0)).max() > 0
    for r in split + [single, plain]:
        r.close()


def test_auto_rule_picks_the_split_on_large_grids_only(plugin_lib, monkeypatch):
    """auto (the default): split when the rectangle holds >= 2^20 cells and >= 3/4 of the grid; and
    at such a size the two forms still agree bit for bit."""
    from mpifdtd_b200.slab import SlabRun
    monkeypatch.delenv("B200FDTD_UNIT_SPLIT", raising=False)
    small = SlabRun("MIE_CYLINDER", "TM_UPML_2D", 256, 256, 8, with_ntff=False)
    assert small.engine.step_form() == 0
    small.close()
    npx, npy, steps = 1100, 2048, 40
    big = SlabRun("ZIGZAG", "TM_UPML_2D", npx, npy, steps, with_ntff=False)
    assert big.engine.step_form() == 1
    for _ in range(steps):
        big.step()
    monkeypatch.setenv("B200FDTD_UNIT_SPLIT", "0")
    plain = SlabRun("ZIGZAG", "TM_UPML_2D", npx, npy, steps, with_ntff=False)
    assert plain.engine.step_form() == 0
    for _ in range(steps):
        plain.step()
    assert np.abs(plain.gather_field(0)).max() > 0
    for slot in range(9):
        assert bit_equal(big.gather_field(slot), plain.gather_field(slot)), slot
    monkeypatch.setenv("B200FDTD_LEAN_INTERIOR", "1")
    lean = SlabRun("ZIGZAG", "TM_UPML_2D", 256, 256, 8, with_ntff=False)
    assert lean.engine.step_form() == 2
    for r in (big, plain, lean):
        r.close()

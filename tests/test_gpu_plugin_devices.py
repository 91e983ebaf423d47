"""Multi-GPU mode of the C plugin (mpifdtd_setDevices / MPIFDTD_DEVICES): ONE process, one host
thread, several y-slab engines with peer halos behind the reference's own entry points
(simulator_init / _calc / _finish, fdtdTM_upml_getEz ...).  Replaces init_mpi + the halo Sendrecv
of mpiTM_UPML.c:196-217,252-334,718-748.  Slab g sits on device g modulo the visible devices, so
one GPU is enough to run it; fields must be bit-identical to the single-engine run, the far-field
files equal to summation order."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import bit_equal, rel_err
from mpifdtd_b200 import binding as B
from test_dropin_cpu import build_driver

pytestmark = pytest.mark.gpu


def gather(gpu, slot):
    """all slabs of a field through the engine handles of the plugin's slabs"""
    L = gpu.L
    L.mpifdtd_upml_slab_engine.restype = C.c_void_p
    L.mpifdtd_upml_slab_engine.argtypes = [C.c_int, C.c_int]
    out = np.zeros((gpu.n_px, gpu.n_py), dtype=np.complex128)
    for g in range(L.mpifdtd_upml_slab_count(gpu.solver)):
        B.check(L.b200fdtd_get_field(C.c_void_p(L.mpifdtd_upml_slab_engine(gpu.solver, g)), slot, out.ctypes.data),
                "get_field")
    return out


@pytest.mark.parametrize("solver,names", [("TM_UPML_2D", ("Ez", "Hx", "Hy")), ("TE_UPML_2D", ("Ex", "Ey", "Hz"))])
@pytest.mark.parametrize("n_slabs,form", [(3, "default"), (4, "one_pass")])
def test_plugin_with_several_slabs_matches_single_engine(plugin_lib, in_tmp_cwd, monkeypatch, solver, names, n_slabs,
                                                         form):
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    if form == "one_pass":
        monkeypatch.setenv("B200FDTD_FUSED", "1")
    npx, npy, steps = 120, 250, 420
    L = plugin_lib
    res = {}
    for n in (1, n_slabs):
        L.mpifdtd_setDevices(n)
        try:
            gpu = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=20)
            assert L.mpifdtd_upml_slab_count(gpu.solver) == n
            gpu.run()
            getters = [gpu.field(f) for f in names]               # the reference's borrowed-pointer getters
            state = [gather(gpu, s) for s in range(9)]
            far = gpu.finish()
        finally:
            L.mpifdtd_setDevices(0)
        res[n] = (getters, state, far)
    assert np.abs(res[1][0][0]).max() > 0
    for a, b in zip(res[n_slabs][0] + res[n_slabs][1], res[1][0] + res[1][1]):
        assert bit_equal(a, b)
    assert rel_err(res[n_slabs][2], res[1][2]) <= 1e-12


@pytest.mark.parametrize("solver,names", [("MPI_TM_UPML_2D", ("Ez", "Hx", "Hy")), ("MPI_TE_UPML_2D", ("Ex", "Ey", "Hz"))])
@pytest.mark.parametrize("n_slabs,form", [(2, "default"), (5, "default"), (3, "unit")])
def test_mpi_variant_solvers_with_several_slabs(plugin_lib, in_tmp_cwd, monkeypatch, solver, names, n_slabs, form):
    """Ids 4/5 -- the reference's own domain-decomposed solvers (E phase first, CW source, all N x N cells
    updated, getters with a ghost ring: mpiTM_UPML.c:196-217,674-743) -- over several slabs: the two
    MPI_Sendrecv exchanges become peer stores.  Fields, state arrays and the projected U/W partial
    sums must reproduce the single-engine run (fields bit for bit, U/W to summation order)."""
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")
    if form == "unit":                              # frame-free rectangle through the unit-coefficient kernels
        monkeypatch.setenv("B200FDTD_UNIT_SPLIT", "1")
    npx, npy, steps = 110, 230, 300
    L = plugin_lib
    res = {}
    for n in (1, n_slabs):
        L.mpifdtd_setDevices(n)
        try:
            gpu = B.Plugin("MIE_CYLINDER", solver, npx, npy, steps=steps, h_u_nm=20, angle_deg=35)
            assert L.mpifdtd_upml_slab_count(gpu.solver) == n
            gpu.run()
            getters = [gpu.field(f).copy() for f in names]
            assert getters[0].shape == (npx + 2, npy + 2)
            state = [gather(gpu, s) for s in range(9)]
            uw = []
            if solver == "MPI_TE_UPML_2D":                      # the series finish() prints (mpiTE_UPML.c:795-878)
                eth = np.zeros((360, steps), dtype=np.complex128)
                eph = np.zeros((360, steps), dtype=np.complex128)
                L.mpifdtd_mpi_te_far_series.argtypes = [C.c_void_p, C.c_void_p]
                assert L.mpifdtd_mpi_te_far_series(eth.ctypes.data, eph.ctypes.data) == 0
                uw = [eth, eph]
            gpu.finish()
        finally:
            L.mpifdtd_setDevices(0)
        res[n] = (getters, state, uw)
    assert np.abs(res[1][0][0]).max() > 0
    for a, b in zip(res[n_slabs][0] + res[n_slabs][1], res[1][0] + res[1][1]):
        assert bit_equal(a, b)
    for a, b in zip(res[n_slabs][2], res[1][2]):
        assert np.abs(b).max() > 0
        assert rel_err(a, b) <= 1e-12


def test_c_driver_with_devices_from_the_environment(plugin_lib, tmp_path):
    """An unmodified C driver of the plugin surface (tests/c/dropin_driver.c) goes multi-GPU by
    environment alone: same stdout line (cell count, peak of the getter's array, material cells),
    same far-field file."""
    exe = build_driver(tmp_path)
    out = {}
    for n in (1, 4):
        work = tmp_path / ("run%d" % n)
        work.mkdir()
        env = dict(os.environ, MPIFDTD_DEVICES=str(n), CUDA_DEVICE_MAX_CONNECTIONS="32", B200FDTD_FUSED="1")
        p = subprocess.run([exe, "128", "300", "2"], capture_output=True, text=True, cwd=work, env=env, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("DRIVER")][0]
        far = np.fromfile(str(work / "0[deg]_380nm_700nm_b.dat")).reshape(321, 360)
        out[n] = (line, far)
    assert out[1][0] == out[4][0] and "peak=0" not in out[1][0]
    assert rel_err(out[4][1], out[1][1]) <= 1e-12

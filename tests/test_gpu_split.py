"""Split-field solvers on the GPU (ids 0 TM, 1 TE: Yee + Berenger PML; 6 NS_TM, 7 NS_TE:
non-standard FDTD) against field snapshots recorded from the unmodified reference
(tests/golden/split_kind*.npz) and, when it travelled, against the reference itself.
Tolerance: fields <= 1e-12 (max-abs-diff / max-abs-ref)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, TOL_FIELD, bit_equal, rel_err
from mpifdtd_b200 import binding as B

pytestmark = pytest.mark.gpu
FIELDS = {0: ["Ez", "Ezx", "Ezy", "Hx", "Hy"], 1: ["Hz", "Hzx", "Hzy", "Ex", "Ey"],
          6: ["Ez", "Ezx", "Ezy", "Hx", "Hy"], 7: ["Hz", "Hzx", "Hzy", "Ex", "Ey"]}
DUMP = {0: "tm_%dnm.txt", 1: "te_%dnm.txt", 6: "ns_tm_%dnm.txt", 7: "ns_te_%dnm.txt"}


@pytest.fixture(params=["lean", "dense"])
def split_form(request, monkeypatch):
    """Kinds 0, 1, 6 run the lean form by default (1-D tables + eps / numerator arrays);
    MPIFDTD_SPLIT_DENSE=1 keeps the reference's eight dense coefficient arrays.  Both must pass."""
    if request.param == "dense":
        monkeypatch.setenv("MPIFDTD_SPLIT_DENSE", "1")
    else:
        monkeypatch.delenv("MPIFDTD_SPLIT_DENSE", raising=False)
    return request.param


@pytest.mark.parametrize("kind,model", [(0, "LAYER"), (0, "MIE_CYLINDER"), (1, "LAYER"), (1, "ZIGZAG"),
                                        (6, "LAYER"), (6, "MORPHO_SCALE")])
def test_lean_form_is_bit_identical_to_dense(plugin_lib, kind, model, in_tmp_cwd, monkeypatch):
    """LAYER puts eps != 1 inside the PML, i.e. the in-kernel field_pmlCoef path with its three
    IEEE divisions; the others cover the 1/eps-only path and the NS numerators."""
    npx, npy, steps = 200, 220, 240      # the validation circle of finish() (1.2 lambda = 76 cells) must fit
    runs = {}
    for form in ("dense", "lean"):
        if form == "dense":
            monkeypatch.setenv("MPIFDTD_SPLIT_DENSE", "1")
        else:
            monkeypatch.delenv("MPIFDTD_SPLIT_DENSE")
        gpu = B.Plugin(model, kind, npx, npy, steps=steps, lambda_nm=633, angle_deg=15)
        bytes_on_device = C_uint64_device_bytes(gpu)
        gpu.run()
        runs[form] = ({f: gpu.field(f) for f in FIELDS[kind]}, bytes_on_device)
        gpu.finish()
    assert np.abs(runs["dense"][0][FIELDS[kind][0]]).max() > 1e-3
    for f in FIELDS[kind]:
        assert bit_equal(runs["lean"][0][f], runs["dense"][0][f]), f
    assert runs["lean"][1] < 0.8 * runs["dense"][1]              # and it really keeps fewer arrays


@pytest.mark.parametrize("model", ["LAYER", "MIE_CYLINDER", "ZIGZAG"])
def test_ns_te_interior_form_is_bit_identical_to_dense(plugin_lib, model, in_tmp_cwd, monkeypatch):
    """Solver id 7 (the reference's own default, main.c:155-156): outside the absorbing frame its
    decay coefficients are exactly 1.0 and C_HZXLX == C_HZYLY, so the kernels read three coefficient
    arrays there instead of eight (b200fdtd_set_split_interior) -- and must give the bits of the
    all-dense form.  LAYER keeps eps != 1 inside the PML (the tanh-dependent frame coefficients).
    (800 columns: two of the four 256-column thread blocks of a row lie wholly inside the rectangle.)"""
    npx, npy, steps = 200, 800, 240
    runs = {}
    for form in ("dense", "interior"):
        if form == "dense":
            monkeypatch.setenv("MPIFDTD_SPLIT_DENSE", "1")
        else:
            monkeypatch.delenv("MPIFDTD_SPLIT_DENSE")
        gpu = B.Plugin(model, 7, npx, npy, steps=steps, lambda_nm=633, angle_deg=15)
        n0 = gpu.launches()
        gpu.step(10)
        per_step = (gpu.launches() - n0) / 10
        gpu.run()
        runs[form] = ({f: gpu.field(f) for f in FIELDS[7]}, per_step)
        gpu.finish()
    assert np.abs(runs["dense"][0][FIELDS[7][0]]).max() > 1e-3
    for f in FIELDS[7]:
        assert bit_equal(runs["interior"][0][f], runs["dense"][0][f]), f
    assert runs["dense"][1] == 2 and runs["interior"][1] == 2       # one launch per phase either way


def C_uint64_device_bytes(gpu):
    import ctypes as C
    n = C.c_uint64(0)
    gpu.L.b200fdtd_device_bytes(gpu.engine_handle(), C.byref(n))
    return n.value


@pytest.mark.parametrize("kind", [0, 1, 6, 7])
def test_split_solver_matches_reference_golden(plugin_lib, kind, in_tmp_cwd, split_form):
    g = np.load(os.path.join(GOLDEN, "split_kind%d.npz" % kind))
    npx, npy, hu, steps, lam, angle = (int(v) for v in g["meta"][:6])
    gpu = B.Plugin("MIE_CYLINDER", kind, npx, npy, steps=steps, h_u_nm=hu, lambda_nm=lam, angle_deg=angle)
    eps_name = {0: "EPS_EZ", 1: "EPS_EY", 6: "EPS_EZ", 7: "EPS_EY"}[kind]
    assert bit_equal(gpu.eps(), g[eps_name])
    gpu.step(steps // 2)
    for f in FIELDS[kind]:
        assert rel_err(gpu.field(f), g["mid_" + f]) <= TOL_FIELD, ("mid", f)
    gpu.step(steps - steps // 2)
    for f in FIELDS[kind]:
        assert rel_err(gpu.field(f), g["end_" + f]) <= TOL_FIELD, ("end", f)
    assert np.abs(gpu.field(FIELDS[kind][0])).max() > 0.1
    gpu.finish()                                   # reset(): validation-circle dump, then free
    lines = open(DUMP[kind] % hu).read().split("\n")
    assert len([l for l in lines if l.strip()]) == 181


@pytest.mark.parametrize("kind,model", [(0, "ZIGZAG"), (1, "LAYER"), (6, "MORPHO_SCALE"), (7, "MIE_CYLINDER")])
def test_split_solver_vs_live_reference(plugin_lib, kind, model, in_tmp_cwd, split_form):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    npx, npy, steps = 200, 220, 260      # the validation circle (1.2 lambda = 76 cells) must fit
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, lambda_nm=633, angle_deg=15)
    ref.run()
    want = {f: ref.field(f) for f in FIELDS[kind]}
    ref.finish()
    os.chdir(cwd)
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, lambda_nm=633, angle_deg=15)
    gpu.run()
    for f in FIELDS[kind]:
        assert rel_err(gpu.field(f), want[f]) <= TOL_FIELD, f
    gpu.finish()


@pytest.mark.parametrize("kind,model,angle", [(0, "LAYER", 20), (1, "MIE_CYLINDER", 0), (6, "ZIGZAG", 15),
                                              (7, "LAYER", 40), (7, "MORPHO_SCALE", 0)])
def test_split_solver_vs_c_oracle(plugin_lib, oracle, kind, model, angle, in_tmp_cwd, split_form):
    """The CUDA path against oracle/split_oracle.c (itself bit-exact against the reference,
    tests/test_oracle_cpu.py), fed with the eps maps the plugin's host code built."""
    import ctypes as C
    npx, npy, steps, lam = 200, 216, 300, 633
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, lambda_nm=lam, angle_deg=angle)
    L = gpu.L
    L.mpifdtd_split_prepare_host(kind)                 # test hook: all three eps maps on the host

    def eps_map(slot):
        buf = (C.c_double * (npx * npy)).from_address(L.mpifdtd_split_dense(kind, 10 + slot))
        return np.frombuffer(buf, dtype=np.float64).reshape(npx, npy).copy()

    cpu = oracle.SplitOracleSim(kind, npx, npy, [eps_map(m) for m in range(3)], lambda_nm=lam, angle_deg=angle)
    gpu.run()
    cpu.step(steps)
    assert np.abs(cpu.field(FIELDS[kind][0])).max() > 1e-3
    for f in FIELDS[kind]:
        assert rel_err(gpu.field(f), cpu.field(f)) <= TOL_FIELD, f
    gpu.finish()
    cpu.close()


# ---------------------------------------------------------------- MPI-variant ids (4, 5)
MPI_FIELDS = {4: ["Ez", "Hx", "Hy", "Jz", "Dz", "Mx", "Bx", "My", "By"],
              5: ["Ex", "Ey", "Hz", "Jx", "Dx", "Jy", "Dy", "Mz", "Bz"]}


@pytest.mark.parametrize("kind,model", [(4, "MIE_CYLINDER"), (5, "MIE_CYLINDER"), (4, "ZIGZAG"), (5, "LAYER")])
def test_mpi_variant_solver_vs_live_reference(plugin_lib, kind, model, in_tmp_cwd):
    """Solver ids 4/5 as rank 0 of 1: E phase first, CW source, all N x N cells updated,
    (N+2) x (N+2) arrays with a ghost ring behind the getters (mpiTM_UPML.c:196-217)."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    npx, npy, steps = 130, 150, 300
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, angle_deg=25)
    ref.run()
    sub = (npx + 2) * (npy + 2)
    want = {f: ref.carray(f, sub).reshape(npx + 2, npy + 2) for f in MPI_FIELDS[kind]}
    eps_name = "EPS_EZ" if kind == 4 else "EPS_EY"
    want_eps = ref.darray(eps_name, sub).reshape(npx + 2, npy + 2)
    os.chdir(cwd)          # (the reference's finish() for these ids calls MPI_Finalize and frees; not needed)
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, angle_deg=25)
    L = gpu.L
    assert (L.mpi_fdtdTM_upml_getSubNpx(), L.mpi_fdtdTM_upml_getSubNcell()) == (npx + 2, sub)
    ptr = L.simulator_getEps()
    import ctypes as C
    mine_eps = np.frombuffer((C.c_double * sub).from_address(ptr), dtype=np.float64).reshape(npx + 2, npy + 2)
    assert bit_equal(mine_eps[1:-1, 1:-1], want_eps[1:-1, 1:-1])
    gpu.run()
    for f in MPI_FIELDS[kind][:3]:
        got = gpu.field(f)
        assert got.shape == (npx + 2, npy + 2)
        assert np.all(got[0, :] == 0) and np.all(got[:, -1] == 0)
        assert rel_err(got, want[f]) <= TOL_FIELD, f
    for slot, f in enumerate(gpu.SLOTS[kind]):
        assert rel_err(gpu.any_field(slot), want[f][1:-1, 1:-1]) <= TOL_FIELD, f
    assert np.abs(want[MPI_FIELDS[kind][0]]).max() > 1e-3
    gpu.finish()


@pytest.mark.parametrize("kind,model,angle", [(4, "MIE_CYLINDER", 25), (5, "MIE_CYLINDER", 25), (4, "ZIGZAG", 0),
                                              (5, "LAYER", 60)])
def test_mpi_variant_solver_vs_oracle(plugin_lib, oracle, kind, model, angle, in_tmp_cwd, monkeypatch):
    """The same ids -- stepping and their own ntff() (row a15) -- against the plain-C restatement
    (oracle_step_mpi, pinned to the reference in tests/test_oracle_cpu.py): needs nothing but this
    repository on the GPU box."""
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")       # all arraySize bins, spills included
    npx, npy, steps = 130, 150, 300
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, angle_deg=angle)
    L = gpu.L
    maps = [(0.0, 0.0, B.D_XY)] if kind == 4 else [(0.5, 0.0, B.D_Y), (0.0, 0.5, B.D_X)]
    eps = []
    for xo, yo, mode in maps:
        e = np.empty((npx, npy))
        L.mpifdtd_fill_eps(e.ctypes.data, xo, yo, mode)
        eps.append(e)
    cpu = oracle.OracleSim(oracle.TM if kind == 4 else oracle.TE, npx, npy, steps, *eps, angle_deg=angle)
    gpu.run()
    cpu.step_mpi(steps)
    for slot, f in enumerate(gpu.SLOTS[kind]):
        want = cpu.field(slot)
        assert np.abs(want).max() > 0, f
        assert rel_err(gpu.any_field(slot), want) <= TOL_FIELD, f
    ring = gpu.field(gpu.SLOTS[kind][0])                  # the getter's array carries the ghost ring
    assert ring.shape == (npx + 2, npy + 2) and rel_err(ring[1:-1, 1:-1], cpu.field(0)) <= TOL_FIELD
    scale = max(np.abs(cpu.uw(slot)).max() for slot in range(3))
    assert scale > 0
    for slot in range(3):                                 # TM Ux, Uy, Wz | TE Wx, Wy, Uz
        got = gpu.ntff_uw(slot, project=(slot == 0))
        assert got.shape == (360, cpu.array_size)
        assert np.abs(got - cpu.uw(slot)).max() <= 1e-10 * scale, slot
    os.makedirs("MPI_TE_UPML", exist_ok=True)
    gpu.finish()
    cpu.close()


UW_NAMES = {4: ["Ux", "Uy", "Wz"], 5: ["Wx", "Wy", "Uz"]}


@pytest.mark.parametrize("kind,model,angle", [(4, "MIE_CYLINDER", 0), (5, "MIE_CYLINDER", 0), (4, "LAYER", 25),
                                              (5, "ZIGZAG", 25)])
def test_mpi_variant_ntff_vs_live_reference(plugin_lib, kind, model, angle, in_tmp_cwd, monkeypatch):
    """ntff() of the MPI-variant solvers (mpiTM_UPML.c:849-1037,
    mpiTE_UPML.c:602-794) -- direct time shift, coef per tap, box indices used as local
    indices -- and, for id 5, the E_theta/E_phi text files finish() writes."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")
    npx, npy, steps = 120, 132, 420
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, angle_deg=angle)
    ref.run()
    want = {n: ref.ntff_uw(n) for n in UW_NAMES[kind]}
    ref_files = {}
    if kind == 5:
        reflib.lib().refhook_mpi_te_upml_ntff_output()
        for stem in ("Eph_r", "Eph_i", "Eth_r", "Eth_i"):
            ref_files[stem] = np.loadtxt(os.path.join(ref.workdir, "MPI_TE_UPML", stem + ".txt"))
    os.chdir(cwd)
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, angle_deg=angle)
    gpu.run()
    scale = max(np.abs(w).max() for w in want.values())
    assert scale > 0
    for slot, n in enumerate(UW_NAMES[kind]):
        got = gpu.ntff_uw(slot, project=(slot == 0))
        assert got.shape == want[n].shape
        assert np.abs(got - want[n]).max() <= 1e-10 * scale, n
    os.makedirs("MPI_TE_UPML", exist_ok=True)
    gpu.finish()
    if kind == 5:
        fscale = max(np.abs(v).max() for v in ref_files.values())
        assert fscale > 0
        for stem, ref_table in ref_files.items():
            mine = np.loadtxt(os.path.join("MPI_TE_UPML", stem + ".txt"))
            assert mine.shape == ref_table.shape == (360, steps)
            assert np.abs(mine - ref_table).max() <= 1e-10 * fscale, stem
    else:
        assert not os.listdir("MPI_TE_UPML")        # id 4 writes nothing (mpiTM_UPML.c:240)


# ---------------------------------------------------------------- opt-in source forms (row a10)
@pytest.mark.parametrize("kind,model,form,hook", [
    (2, "MIE_CYLINDER", "CW", "refhook_tm_upml_update_cw"),
    (2, "NO_MODEL", "PLANE", "refhook_tm_upml_update_plane_wave"),
    (4, "NO_MODEL", "PLANE", "refhook_mpi_tm_upml_update_plane_wave"),
    (4, "ZIGZAG", "PLANE", "refhook_mpi_tm_upml_update_plane_wave")])
def test_optional_source_forms_vs_live_reference(plugin_lib, kind, model, form, hook, in_tmp_cwd, monkeypatch):
    """field_scatteredWave (the commented alternative at fdtdTM_upml.c:62) and planeWave
    (mpiTM_UPML.c:377-403, commented call at :204): the reference's own static functions,
    composed by oracle/refbuild/wrap_*.c in the order the commented lines give."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")
    npx, npy, steps = (256, 256, 300) if model == "NO_MODEL" else (120, 140, 360)
    cwd = os.getcwd()
    ref = reflib.RefSim(model, kind, npx, npy, steps=steps, angle_deg=20)
    ref.step_fn(hook, steps)
    ring = 2 if kind == 4 else 0
    shape = (npx + ring, npy + ring)
    want = {f: ref.carray(f, shape[0] * shape[1]).reshape(shape) for f in ("Ez", "Hx", "Hy")}
    want_uw = {n: ref.ntff_uw(n) for n in ("Ux", "Uy", "Wz")}
    os.chdir(cwd)
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, angle_deg=20, source_form=form)
    try:
        gpu.run()
        assert np.abs(want["Ez"]).max() > 1e-3
        for f in ("Ez", "Hx", "Hy"):
            assert rel_err(gpu.field(f), want[f]) <= TOL_FIELD, f
        scale = max(np.abs(w).max() for w in want_uw.values())
        for slot, n in enumerate(("Ux", "Uy", "Wz")):
            got = gpu.ntff_uw(slot, project=(slot == 0))
            d = np.abs(got - want_uw[n])
            assert d.max() <= 1e-10 * scale, (n, np.unravel_index(d.argmax(), d.shape), d.max())
        gpu.finish()
    finally:
        B.lib().mpifdtd_setSourceForm(0)


@pytest.mark.parametrize("kind,model,form", [(2, "MIE_CYLINDER", "CW"), (2, "NO_MODEL", "PLANE"),
                                             (4, "NO_MODEL", "PLANE"), (4, "ZIGZAG", "PLANE")])
def test_optional_source_forms_vs_oracle(plugin_lib, oracle, kind, model, form, in_tmp_cwd, monkeypatch):
    """The same forms against the plain-C restatement (pinned to the reference's functions in
    tests/test_oracle_cpu.py::test_optional_source_restatements_vs_live_reference)."""
    monkeypatch.setenv("MPIFDTD_NTFF_FULL_BINS", "1")
    npx, npy, steps = (256, 256, 300) if model == "NO_MODEL" else (120, 140, 360)
    gpu = B.Plugin(model, kind, npx, npy, steps=steps, angle_deg=20, source_form=form)
    try:
        eps = np.empty((npx, npy))
        gpu.L.mpifdtd_fill_eps(eps.ctypes.data, 0.0, 0.0, B.D_XY)
        cpu = oracle.OracleSim(oracle.TM, npx, npy, steps, eps, angle_deg=20, source_form=form)
        gpu.run()
        if kind == 4:
            cpu.step_mpi(steps)
        else:
            cpu.step(steps)
        assert np.abs(cpu.field(0)).max() > 1e-3
        for slot in (0, 3, 6):                            # Ez, Hx, Hy
            assert rel_err(gpu.any_field(slot), cpu.field(slot)) <= TOL_FIELD, slot
        scale = max(np.abs(cpu.uw(slot)).max() for slot in range(3))
        for slot in range(3):
            got = gpu.ntff_uw(slot, project=(slot == 0))
            assert np.abs(got - cpu.uw(slot)).max() <= 1e-10 * scale, slot
        os.makedirs("MPI_TE_UPML", exist_ok=True)
        gpu.finish()
        cpu.close()
    finally:
        B.lib().mpifdtd_setSourceForm(0)


# ---------------------------------------------------------------- frequency-domain NTFF
def test_frequency_ntff_tm_upml_vs_oracle(plugin_lib, oracle, in_tmp_cwd):
    """ntffTM_Frequency (ntffTM.c:72-158) on the GPU fields of the serial TM UPML solver."""
    n, steps = 120, 520
    gpu = B.Plugin("MIE_CYLINDER", "TM_UPML_2D", n, steps=steps, h_u_nm=20)
    cpu = oracle.OracleSim(oracle.TM, n, n, steps, gpu.eps(), h_u_nm=20)
    gpu.run()
    cpu.step(steps)
    mine = np.zeros(360, dtype=np.complex128)
    assert gpu.L.mpifdtd_ntffFrequency(2, mine.ctypes.data) == 0
    want = cpu.frequency_tm()
    assert np.abs(want).max() > 0
    assert rel_err(mine, want) <= 1e-10
    gpu.finish()


def test_frequency_ntff_ns_tm_vs_live_reference(plugin_lib, in_tmp_cwd):
    """BASELINE configs[3] in miniature: NS-FDTD TM, frequency-domain far field of the final
    fields (the reference's NS solver has no NTFF; its ntffTM_Frequency is applied to its
    fields, after ntffTM_init() for R0, as SURVEY 8d prescribes)."""
    import ctypes as C
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    npx, npy, steps, lam = 200, 220, 320, 600
    cwd = os.getcwd()
    ref = reflib.RefSim("MIE_CYLINDER", 6, npx, npy, steps=steps, lambda_nm=lam)
    L = reflib.lib()
    L.ntffTM_init()
    ref.run()
    want = np.zeros(360, dtype=np.complex128)
    L.ntffTM_Frequency.argtypes = [C.c_void_p] * 4
    L.ntffTM_Frequency(ref._ptr("Hx"), ref._ptr("Hy"), ref._ptr("Ez"), want.ctypes.data)
    ref.finish()
    os.chdir(cwd)
    gpu = B.Plugin("MIE_CYLINDER", 6, npx, npy, steps=steps, lambda_nm=lam)
    gpu.run()
    mine = np.zeros(360, dtype=np.complex128)
    assert gpu.L.mpifdtd_ntffFrequency(6, mine.ctypes.data) == 0
    assert np.abs(want).max() > 0 and rel_err(mine, want) <= 1e-10
    gpu.finish()

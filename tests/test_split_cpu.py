"""Host-built dense coefficient arrays, permittivity maps and source factors of the four
split-field solvers (ids 0, 1, 6, 7), bit-exact against the arrays of the unmodified
reference (fdtdTM.c:197-242, fdtdTE.c:199-243, nsFdtdTM.c:231-307, nsFdtdTE.c:100-181).
Needs oracle/_ref/libref.so; golden copies for the GPU box live in tests/golden/split_*.npz."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN, bit_equal
from mpifdtd_b200 import binding as B

NAMES = {0: ["C_EZX", "C_EZXLX", "C_EZY", "C_EZYLY", "C_HX", "C_HXLY", "C_HY", "C_HYLX"],
         1: ["C_EX", "C_EXLY", "C_EY", "C_EYLX", "C_HZX", "C_HZXLX", "C_HZY", "C_HZYLY"],
         6: ["C_EZX", "C_EZXLX", "C_EZY", "C_EZYLY", "C_HX", "C_HXLY", "C_HY", "C_HYLX"],
         7: ["C_EX", "C_EXLY", "C_EY", "C_EYLX", "C_HZX", "C_HZXLX", "C_HZY", "C_HZYLY"]}
EPS = {0: ["EPS_EZ", "EPS_HX", "EPS_HY"], 1: ["EPS_EX", "EPS_EY", "EPS_HZ"],
       6: ["EPS_EZ", "EPS_HX", "EPS_HY"], 7: ["EPS_EX", "EPS_EY", "EPS_HZ"]}


def product_arrays(L, model, kind, npx, npy, hu, lam):
    L.models_setModel(B.MODELS[model])
    L.field_init(B.FieldInfo(npx * hu, npy * hu, hu, 10, lam, 0, 10))
    L.models_initModel()
    L.mpifdtd_split_prepare_host(kind)
    out = {}
    for slot, name in enumerate(NAMES[kind] + ["SRC0", "SRC1"] + EPS[kind]):
        ptr = L.mpifdtd_split_dense(kind, slot)
        buf = (C.c_double * (npx * npy)).from_address(ptr)
        out[name] = np.frombuffer(buf, dtype=np.float64).reshape(npx, npy).copy()
    return out


@pytest.mark.parametrize("kind", [0, 1, 6, 7])
@pytest.mark.parametrize("model,npx,npy,hu,lam", [("MIE_CYLINDER", 96, 104, 20, 500), ("ZIGZAG", 80, 150, 10, 633)])
def test_dense_arrays_bit_exact_vs_reference(plugin_lib, kind, model, npx, npy, hu, lam):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    ref = reflib.RefSim(model, kind, npx, npy, steps=10, h_u_nm=hu, lambda_nm=lam)
    want = {n: ref.coef(n) for n in NAMES[kind] + EPS[kind]}
    ref.finish()
    mine = product_arrays(plugin_lib, model, kind, npx, npy, hu, lam)
    for n in NAMES[kind] + EPS[kind]:
        assert bit_equal(mine[n], want[n]), (kind, n, int((mine[n] != want[n]).sum()))


@pytest.mark.parametrize("kind", [0, 1, 6, 7])
def test_dense_arrays_bit_exact_vs_golden(plugin_lib, kind):
    g = np.load(os.path.join(GOLDEN, "split_kind%d.npz" % kind))
    npx, npy, hu, steps, lam = (int(v) for v in g["meta"][:5])
    mine = product_arrays(plugin_lib, "MIE_CYLINDER", kind, npx, npy, hu, lam)
    for n in NAMES[kind] + EPS[kind]:
        assert bit_equal(mine[n], g[n]), (kind, n)
    # vacuum cells never receive a source: the factor is exactly zero there
    eps_src = mine[EPS[kind][0]] if kind in (0, 6) else mine[EPS[kind][1]]
    src = mine["SRC0"] if kind in (0, 6) else mine["SRC1"]
    interior = np.zeros_like(src, dtype=bool)
    interior[1:-1, 1:-1] = True
    assert np.all(src[(eps_src == 1.0) & interior] == 0.0)
    assert np.any(src != 0.0)


# ---------------------------------------------------------------- lean form, on the host
def _pml_pair(eps, sig):
    """field_pmlCoef / field_pmlCoef_LXY as split_kernels.cu evaluates them (IEEE double ops)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(eps == 1.0, 1.0, 1.0 / eps)
        q = sig / eps
        c = np.where(sig != 0.0, (1.0 - q) / (1.0 + q), 1.0)
        l = np.where(sig != 0.0, 1.0 / (eps + sig), inv)
    return c, l, inv


@pytest.mark.parametrize("kind", [0, 1, 6])
@pytest.mark.parametrize("model,npx,npy,hu,lam", [("LAYER", 90, 130, 10, 500), ("MIE_CYLINDER", 96, 104, 20, 633)])
def test_lean_tables_expand_to_the_reference_dense_arrays(plugin_lib, kind, model, npx, npy, hu, lam):
    """The lean form keeps 1-D tables + eps (kinds 0, 1) or numerator arrays (kind 6) and forms
    the coefficients in the kernel.  The same formation in numpy (IEEE double) must reproduce
    the reference's eight dense arrays and source factors bit for bit -- LAYER puts eps != 1
    inside the PML, where the three-division path runs."""
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not present")
    L = plugin_lib
    ref = reflib.RefSim(model, kind, npx, npy, steps=10, h_u_nm=hu, lambda_nm=lam)
    want = {n: ref.coef(n) for n in NAMES[kind]}
    ref.finish()
    dense = product_arrays(L, model, kind, npx, npy, hu, lam)       # also leaves field_init state for this grid
    L.mpifdtd_split_prepare_host_lean(kind)
    ti, tj = np.empty((4, npx)), np.empty((4, npy))
    L.mpifdtd_split_lean_tables(kind, ti.ctypes.data, tj.ctypes.data)

    def host(slot):
        buf = (C.c_double * (npx * npy)).from_address(L.mpifdtd_split_dense(kind, slot))
        return np.frombuffer(buf, dtype=np.float64).reshape(npx, npy).copy()

    I, J = np.arange(npx)[:, None], np.arange(npy)[None, :]
    by_i = lambda slot: np.broadcast_to(ti[slot][:, None], (npx, npy))
    by_j = lambda slot: np.broadcast_to(tj[slot][None, :], (npx, npy))
    got = {}
    if kind == 0:
        eps = host(10)
        got["C_EZX"], got["C_EZXLX"], inv = _pml_pair(eps, by_i(0))
        got["C_EZY"], got["C_EZYLY"], _ = _pml_pair(eps, by_j(0))
        got["C_HY"], got["C_HYLX"], got["C_HX"], got["C_HXLY"] = by_i(1), by_i(2), by_j(1), by_j(2)
        assert bit_equal(inv - 1.0, dense["SRC0"])
    elif kind == 1:
        eps_x, eps_y = host(10), host(11)
        got["C_EX"], got["C_EXLY"], _ = _pml_pair(eps_x, by_j(0))
        got["C_EY"], got["C_EYLX"], inv_y = _pml_pair(eps_y, by_i(0))
        got["C_HZX"], got["C_HZXLX"], got["C_HZY"], got["C_HZYLY"] = by_i(1), by_i(2), by_j(1), by_j(2)
        assert bit_equal(inv_y - 1.0, dense["SRC1"])
    else:
        over = lambda g, den: np.where(den == 1.0, g, g / den)
        got["C_EZX"], got["C_EZY"] = by_i(0), by_j(0)
        got["C_EZXLX"] = got["C_EZYLY"] = over(host(1), by_i(1))
        got["C_HX"], got["C_HY"] = by_j(1), by_i(2)
        got["C_HXLY"], got["C_HYLX"] = over(host(5), by_j(2)), over(host(7), by_i(3))
        assert bit_equal(host(8), dense["SRC0"])
    inner = (slice(1, -1), slice(1, -1)) if kind == 7 else (slice(None), slice(None))
    for n in NAMES[kind]:
        assert bit_equal(np.ascontiguousarray(got[n][inner]), np.ascontiguousarray(want[n][inner])), (kind, n)
    assert (I + J).shape == (npx, npy)

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

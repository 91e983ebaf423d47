"""The twelve 1-D UPML tables the plugin hands to the GPU engine, expanded back to
the reference's dense N_CELL coefficient arrays and compared bit-exactly with
arrays recorded from the reference (fdtdTM_upml.c:224-274, fdtdTE_upml.c:361-412)."""
import numpy as np
import pytest

from helpers import bit_equal, golden
from mpifdtd_b200 import binding as B

TM = ["C_JZ", "C_JZHXHY", "C_DZ", "C_DZJZ1", "C_DZJZ0", "C_MX", "C_MXEZ", "C_BX", "C_BXMX1", "C_BXMX0",
      "C_MY", "C_MYEZ", "C_BY", "C_BYMY1", "C_BYMY0"]
TE = ["C_JX", "C_JXHZ", "C_DX", "C_DXJX1", "C_DXJX0", "C_JY", "C_JYHZ", "C_DY", "C_DYJY1", "C_DYJY0",
      "C_MZ", "C_MZEXEY", "C_BZ", "C_BZMZ1", "C_BZMZ0"]


@pytest.mark.parametrize("fixture,kind,names", [("mie_tm_upml_88x96", 2, TM), ("mie_te_upml_88x96", 3, TE)])
def test_dense_coefficients_bit_exact(plugin_lib, fixture, kind, names):
    g = golden(fixture + ".npz")
    npx, npy, hu, steps = (int(v) for v in g["meta"][:4])
    L = plugin_lib
    L.models_setModel(B.MODELS["MIE_CYLINDER"])
    L.field_init(B.FieldInfo(npx * hu, npy * hu, hu, 10, 500, 0, steps))
    for name in names:
        dense = np.empty((npx, npy))
        assert L.mpifdtd_upml_dense_coefficient(kind, name.encode(), dense.ctypes.data) == 0, name
        assert bit_equal(dense, g[name]), name


def test_interior_coefficients_are_exactly_one(plugin_lib):
    """Outside the PML every sigma is 0, so all recurrences reduce to J += curl etc."""
    L = plugin_lib
    L.field_init(B.FieldInfo(640, 700, 10, 10, 500, 0, 10))
    ti, tj = np.empty((6, 64)), np.empty((6, 70))
    for kind in (2, 3):
        L.mpifdtd_upml_tables(kind, ti.ctypes.data, tj.ctypes.data)
        num_den = {2: ([4, 5], [5]), 3: ([2, 3], [3])}[kind]     # NUM_* (by j) and DEN_* (by i) slots
        for s in range(6):
            want_i = 2.0 if s in num_den[1] else 1.0
            want_j = 2.0 if s in num_den[0] else 1.0
            assert np.all(ti[s, 10:53] == want_i), (kind, "i", s)
            assert np.all(tj[s, 10:59] == want_j), (kind, "j", s)
        assert np.any(ti[:, :10] != ti[:, 20:30]) and np.any(tj[:, 60:] != tj[:, 20:30])


@pytest.mark.parametrize("kind", [2, 3, 4, 5])
def test_frame_free_region_from_the_tables(plugin_lib, kind):
    """b200fdtd_upml_interior (what B200FDTD_OPT_LEAN_INTERIOR relies on): inside the region every
    dense coefficient of the reference is exactly 1, and the region cannot be grown."""
    L = plugin_lib
    npx, npy = 64, 70
    L.models_setModel(B.MODELS["MIE_CYLINDER"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, 10, 500, 0, 10))
    ti, tj = np.empty((6, npx)), np.empty((6, npy))
    L.mpifdtd_upml_tables(kind, ti.ctypes.data, tj.ctypes.data)
    out = np.zeros(4, dtype=np.int32)
    assert L.b200fdtd_upml_interior(kind, ti.ctypes.data, npx, tj.ctypes.data, npy, out.ctypes.data) == 0
    i_lo, i_hi, j_lo, j_hi = (int(v) for v in out)
    assert (i_lo, i_hi, j_lo, j_hi) == (10, npx - 11, 10, npy - 11)
    if kind in (2, 3):        # the dense-coefficient probe serves the serial kinds
        grown = np.zeros((npx, npy), dtype=bool)
        for name in (TM if kind == 2 else TE):
            dense = np.empty((npx, npy))
            assert L.mpifdtd_upml_dense_coefficient(kind, name.encode(), dense.ctypes.data) == 0, name
            assert np.all(dense[i_lo:i_hi + 1, j_lo:j_hi + 1] == 1.0), name
            grown |= dense != 1.0
        # one more row or column on any side meets a coefficient != 1
        assert grown[i_lo - 1, j_lo:j_hi + 1].any() and grown[i_hi + 1, j_lo:j_hi + 1].any()
        assert grown[i_lo:i_hi + 1, j_lo - 1].any() and grown[i_lo:i_hi + 1, j_hi + 1].any()
    else:
        assert len(np.unique(ti[:, i_lo:i_hi + 1])) <= 2 and len(np.unique(tj[:, j_lo:j_hi + 1])) <= 2
        assert len(np.unique(ti[:, i_lo - 1:i_hi + 2])) > 2 and len(np.unique(tj[:, j_lo - 1:j_hi + 2])) > 2
    # a grid that is all frame has no such region
    L.field_init(B.FieldInfo(200, 200, 10, 10, 500, 0, 10))
    ti, tj = np.empty((6, 20)), np.empty((6, 20))
    L.mpifdtd_upml_tables(kind, ti.ctypes.data, tj.ctypes.data)
    assert L.b200fdtd_upml_interior(kind, ti.ctypes.data, 20, tj.ctypes.data, 20, out.ctypes.data) == 0
    assert out[0] > out[1] and out[2] > out[3]


@pytest.mark.parametrize("npx,npy,pml", [(40, 52, 5), (100, 64, 16), (33, 47, 10), (256, 300, 12)])
@pytest.mark.parametrize("kind", [2, 3])
def test_frame_free_region_follows_the_pml_width(plugin_lib, kind, npx, npy, pml):
    """field_sigmaX/Y vanish for pml <= u < N - pml - 1 at the integer and the half-cell position
    (field.c:259-283), so the rectangle is [pml, N - pml - 1] in both directions."""
    L = plugin_lib
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(npx * 10, npy * 10, 10, pml, 500, 0, 10))
    ti, tj = np.empty((6, npx)), np.empty((6, npy))
    L.mpifdtd_upml_tables(kind, ti.ctypes.data, tj.ctypes.data)
    out = np.zeros(4, dtype=np.int32)
    assert L.b200fdtd_upml_interior(kind, ti.ctypes.data, npx, tj.ctypes.data, npy, out.ctypes.data) == 0
    assert tuple(int(v) for v in out) == (pml, npx - pml - 1, pml, npy - pml - 1)

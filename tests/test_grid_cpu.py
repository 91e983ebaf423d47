"""field_init / sigma / soft-start clock of the plugin's host layer against values
recorded from the unmodified reference (tests/golden/grid_cases.json).
Reference: field.c:89-143, 259-283, 312-315."""
import ctypes as C

import pytest

from helpers import golden_json
from mpifdtd_b200 import binding as B

CASES = golden_json("grid_cases.json")


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_hu%d" % (c["width_nm"], c["height_nm"], c["h_u_nm"]))
def test_grid_and_ntff_box_bit_exact(plugin_lib, case):
    L = plugin_lib
    L.field_init(B.FieldInfo(case["width_nm"], case["height_nm"], case["h_u_nm"], case["pml"],
                             case["lambda_nm"], 0, case["steps"]))
    g = L.field_getFieldInfo_S()
    assert (g.N_PX, g.N_PY, g.N_PML) == (case["N_PX"], case["N_PY"], case["pml"])
    assert g.N_X == case["N_PX"] - 2 * case["pml"] and g.DX == case["N_PY"] and g.DY == 1
    assert C.c_int.in_dll(L, "N_PX").value == case["N_PX"]
    b = L.field_getNTFFInfo()
    for key in ("top", "bottom", "left", "right", "cx", "cy", "arraySize"):
        assert getattr(b, key) == case[key], key
    assert b.RFperC == case["RFperC"]
    assert L.field_getOmega().hex() == case["omega"]
    assert L.field_getK().hex() == case["k"]
    for x, want in case["sigma_x"]:
        assert L.field_sigmaX(x, 0.0).hex() == want, ("sigmaX", x)
    for y, want in case["sigma_y"]:
        assert L.field_sigmaY(0.0, y).hex() == want, ("sigmaY", y)
    L.field_reset()
    for want in case["ray_coef"]:
        L.field_nextStep()
        assert L.field_getRayCoef().hex() == want
    assert L.field_getTime() == 5.0


def test_time_and_finish_flags(plugin_lib):
    L = plugin_lib
    L.field_init(B.FieldInfo(640, 640, 10, 10, 500, 0, 3))
    assert L.field_getTime() == 0.0 and not L.field_isFinish()
    for _ in range(3):
        L.field_nextStep()
    assert L.field_isFinish()
    L.field_reset()
    assert L.field_getTime() == 0.0 and L.field_getRayCoef() == 0.0
    L.field_setWaveAngle(35)
    assert L.field_getWaveAngle() == 35.0
    assert L.field_index(3, 4) == 3 * 64 + 4 and L.ind(2, 1) == 2 * 64 + 1

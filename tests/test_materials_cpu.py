"""Permittivity maps of the plugin's material models, bit-exact against maps the
unmodified reference produced (tests/golden/eps_maps.npz; models.c:132-150 and the
seven model files).  When oracle/_ref/libref.so is present the same check also
runs live against the reference on extra grid shapes."""
import os
import shutil

import numpy as np
import pytest

from helpers import GOLDEN, bit_equal, golden
from mpifdtd_b200 import binding as B

MODES = {"ez": (0.0, 0.0, B.D_XY), "ex": (0.5, 0.0, B.D_Y), "ey": (0.0, 0.5, B.D_X),
         "hzx": (0.5, 0.5, B.D_X)}
EPS = golden("eps_maps.npz")
TRACE_IMAGE_SRC = os.path.join(GOLDEN, "traceImage_fixture.txt")


def product_eps(L, model, npx, npy, hu, xo, yo, mode, pml=10):
    L.models_setModel(B.MODELS[model])
    L.field_init(B.FieldInfo(npx * hu, npy * hu, hu, pml, 500, 0, 10))
    L.models_initModel()
    out = np.empty((npx, npy))
    L.mpifdtd_fill_eps(out.ctypes.data, xo, yo, mode)
    return out


def parse_key(key):
    model, dims, hu, tag = key.rsplit("_", 3)
    npx, npy = (int(v) for v in dims.split("x"))
    return model, npx, npy, int(hu[2:]), tag


@pytest.mark.parametrize("key", sorted(EPS.files))
def test_eps_map_bit_exact_vs_golden(plugin_lib, key, in_tmp_cwd):
    model, npx, npy, hu, tag = parse_key(key)
    if model == "TRACE_IMAGE":
        # upstream opens "traceImage1.txt" in cwd (traceImageModel.c:8,82)
        shutil.copy(TRACE_IMAGE_SRC, "traceImage1.txt")
    xo, yo, mode = MODES[tag]
    mine = product_eps(plugin_lib, model, npx, npy, hu, xo, yo, mode)
    assert bit_equal(mine, EPS[key]), "%s: %d cells differ" % (key, int((mine != EPS[key]).sum()))


def test_slab_fill_equals_global_columns(plugin_lib):
    L = plugin_lib
    full = product_eps(L, "ZIGZAG", 90, 150, 10, 0.0, 0.0, B.D_XY)
    for j0, nj in [(0, 75), (75, 75), (40, 61), (149, 1)]:
        slab = np.empty((90, nj))
        L.mpifdtd_fill_eps_slab(slab.ctypes.data, 0.0, 0.0, B.D_XY, j0, nj)
        assert bit_equal(slab, full[:, j0:j0 + nj])


def test_model_iterators_and_sizes(plugin_lib):
    """needSize / isFinish sweeps (circleModel.c:60-88, zigzagModel.c:80-144,
    multiLayerModel.c:235-283, morphoScaleModel.c:256-285,435-440)."""
    import ctypes as C
    L = plugin_lib
    x, y = C.c_int(), C.c_int()
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.models_needSize(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (1000, 1000) and L.models_isFinish()
    L.models_setModel(B.MODELS["MIE_CYLINDER"])
    L.models_needSize(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (2300, 2300)
    L.models_setModel(B.MODELS["ZIGZAG"])
    L.models_needSize(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (130, 1560)      # cos/sin(80 deg) * 300, 5 layers, thick 80
    L.models_setModel(B.MODELS["LAYER"])
    L.models_needSize(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (300, 760)       # width 300; (100+90)*4
    L.models_setModel(B.MODELS["MORPHO_SCALE"])
    L.models_needSize(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (300, 1080)      # width 300; (90+90)*6


def test_concentric_is_disabled_like_upstream(tmp_path):
    """models.c:82-90: selecting the concentric model prints and exit(2)s."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from mpifdtd_b200 import binding as B; "
            "B.lib().models_setModel(4)" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ)
    env.pop("MPIFDTD_ENABLE_CONCENTRIC", None)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode == 2 and "not implemented concentricCircle Model" in p.stdout


reflib_available = os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libref.so")) \
    and os.path.exists("/root/reference")


@pytest.mark.skipif(not reflib_available, reason="needs oracle/_ref/libref.so and /root/reference")
@pytest.mark.parametrize("model", ["MIE_CYLINDER", "LAYER", "MORPHO_SCALE", "ZIGZAG"])
def test_eps_map_bit_exact_vs_live_reference(plugin_lib, model):
    from oracle import reflib
    for (npx, npy, hu) in [(70, 131, 10), (121, 97, 5)]:
        for tag in ("ez", "ex", "ey"):
            xo, yo, mode = MODES[tag]
            want = reflib.eps_map(model, npx, npy, xo, yo, mode, h_u_nm=hu)
            mine = product_eps(plugin_lib, model, npx, npy, hu, xo, yo, mode)
            assert bit_equal(mine, want)


def test_eps_palette_round_trips_every_model(plugin_lib):
    """mpifdtd_eps_palette: 16-bit indices into the table of a map's distinct values reproduce the map
    bit for bit (what b200fdtd_set_eps_palette ships over PCIe instead of the doubles); a map with more
    than 65536 distinct values is refused (-1: the dense upload is used)."""
    import ctypes as C
    L = plugin_lib
    n = 160
    for model in ("MIE_CYLINDER", "LAYER", "ZIGZAG", "MORPHO_SCALE", "NO_MODEL"):
        L.models_setModel(B.MODELS[model])
        L.field_init(B.FieldInfo(n * 10, n * 10, 10, 10, 500, 0, 10))
        L.models_initModel()
        eps = np.empty((n, n))
        L.mpifdtd_fill_eps(eps.ctypes.data, 0.0, 0.0, B.D_XY)
        index, table = np.zeros((n, n), dtype=np.uint16), np.zeros(65536)
        count = L.mpifdtd_eps_palette(eps.ctypes.data, eps.size, index.ctypes.data, table.ctypes.data)
        assert 1 <= count <= 65536 and count == len(np.unique(eps)), model
        assert np.array_equal(table[index].view(np.uint64), eps.view(np.uint64)), model
    rich = np.random.default_rng(0).random((300, 300))
    index, table = np.zeros(rich.shape, dtype=np.uint16), np.zeros(65536)
    assert L.mpifdtd_eps_palette(rich.ctypes.data, rich.size, index.ctypes.data, table.ctypes.data) == -1

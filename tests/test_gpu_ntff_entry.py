"""The reference's public NTFF entry points (ntffTM.h:7-28, ntffTE.h:5-15) exported for a solver
file that keeps its fields on the host: same signatures, accumulation on the GPU.  Checked against
the reference's OWN functions (oracle/_ref/libref.so) fed with the same host arrays step by step:
U / W over all arraySize bins, E_theta / E_phi, the far-field file, ntffTM_Frequency."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import TOL_FARFIELD, rel_err
from mpifdtd_b200 import binding as B

pytestmark = pytest.mark.gpu
PAD = 64


def padded(n):
    """complex array with slack on both sides: upstream's calc() writes a few bins outside its rows
    (ntffTM.c:285-287, the row-spill quirk), at the very ends outside the allocation"""
    buf = np.zeros(n + 2 * PAD, dtype=np.complex128)
    return buf, buf[PAD:PAD + n]


@pytest.mark.parametrize("tm", [True, False])
def test_entry_points_vs_the_reference_functions(plugin_lib, tm, tmp_path):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so did not travel with this snapshot")
    R, L = reflib.lib(), plugin_lib
    npx, npy, steps = 84, 100, 36
    info = (npx * 10, npy * 10, 10, 10, 500, 20, steps)
    L.models_setModel(B.MODELS["NO_MODEL"])
    L.field_init(B.FieldInfo(*info))
    R.models_setModel(0)
    R.field_init(reflib.FieldInfo(*info))
    box = L.field_getNTFFInfo()
    n_acc = 360 * box.arraySize
    pre = "ntffTM" if tm else "ntffTE"
    for lib in (L, R):
        getattr(lib, pre + "_init")()
    rng = np.random.default_rng(5 + tm)
    fields = [rng.standard_normal(npx * npy) + 1j * rng.standard_normal(npx * npy) for _ in range(3)]
    acc = {"ours": [padded(n_acc) for _ in range(3)], "ref": [padded(n_acc) for _ in range(3)]}
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for lib in (L, R):
        lib.field_reset()
    for step in range(steps):
        f = [x * (1.0 + 0.05 * step) * np.exp(0.3j * step) for x in fields]
        for name, lib in (("ours", L), ("ref", R)):
            getattr(lib, pre + "_TimeCalc")(ptr(f[0]), ptr(f[1]), ptr(f[2]), *[ptr(a[1]) for a in acc[name]])
            lib.field_nextStep()
    # E_theta / E_phi: the call also stores the accumulators into the caller's arrays
    out = {}
    for name, lib in (("ours", L), ("ref", R)):
        eth, eph = np.zeros(n_acc, dtype=np.complex128), np.zeros(n_acc, dtype=np.complex128)
        getattr(lib, pre + "_TimeTranslate")(*[ptr(a[1]) for a in acc[name]], ptr(eth), ptr(eph))
        out[name] = (eth, eph)
    for m in range(3):
        want = acc["ref"][m][1]
        assert np.abs(want).max() > 0
        assert rel_err(acc["ours"][m][1], want) <= 1e-13, m
    for m in range(2):
        assert rel_err(out["ours"][m], out["ref"][m]) <= 1e-13, m
    # the far-field file
    tables = {}
    for name, lib in (("ours", L), ("ref", R)):
        work = tmp_path / name
        work.mkdir()
        cwd = os.getcwd()
        os.chdir(work)
        try:
            getattr(lib, pre + "_TimeOutput")(*[ptr(a[1]) for a in acc[name]])
        finally:
            os.chdir(cwd)
        tables[name] = np.fromfile(str(work / "20[deg]_380nm_700nm_b.dat")).reshape(321, 360)
    assert tables["ref"].max() > 0 and rel_err(tables["ours"], tables["ref"]) <= TOL_FARFIELD
    if tm:
        res = {}
        for name, lib in (("ours", L), ("ref", R)):
            r = np.zeros(360, dtype=np.complex128)
            lib.ntffTM_Frequency(ptr(fields[0]), ptr(fields[1]), ptr(fields[2]), ptr(r))
            res[name] = r
        assert np.abs(res["ref"]).max() > 0 and rel_err(res["ours"], res["ref"]) <= 1e-14
    getattr(L, pre + "_finish")()

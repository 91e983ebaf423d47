"""On-disk far-field formats (ntff.c:6-33): the plugin's writers produce the
same bytes as the reference's own ntff_outputEnormTxt / ntff_outputEnormBin for the same table."""
import ctypes as C
import os

import numpy as np
import pytest


def table_pointers(table):
    rows = (C.POINTER(C.c_double) * table.shape[0])()
    for r in range(table.shape[0]):
        rows[r] = table[r].ctypes.data_as(C.POINTER(C.c_double))
    return rows


@pytest.mark.parametrize("seed", [0, 1])
def test_far_field_writers_match_the_reference_byte_for_byte(plugin_lib, tmp_path, seed):
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libref.so not built here")
    rng = np.random.default_rng(seed)
    table = np.ascontiguousarray(rng.random((321, 360)) * 10.0 ** rng.integers(-9, 4, (321, 360)))
    table[5, 7] = 0.0
    table[6, 8] = -1.25e-7
    table[7, 9] = 123456789.123456789
    table[8, 10] = 0.0000005           # rounding knife-edge of "%lf"
    if seed == 1:
        table[9, 11] = 3.0e25          # beyond the bounded fast path
    rows = table_pointers(table)
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref = reflib.lib()
        for lib, tag in ((plugin_lib, "mine"), (ref, "ref")):
            lib.ntff_outputEnormTxt.argtypes = [C.c_void_p, C.c_char_p]
            lib.ntff_outputEnormBin.argtypes = [C.c_void_p, C.c_char_p]
            lib.ntff_outputEnormTxt(rows, (tag + ".txt").encode())
            lib.ntff_outputEnormBin(rows, (tag + ".dat").encode())
        assert open("mine.txt", "rb").read() == open("ref.txt", "rb").read()
        assert open("mine.dat", "rb").read() == open("ref.dat", "rb").read()
        assert os.path.getsize("mine.dat") == 924480
    finally:
        os.chdir(old)

"""Launch geometry of the split step forms (unit-coefficient / lean interior kernel + frame kernel
over a rectangle table, upml_kernels.cu): interior launch and frame launch together must visit every
updated cell exactly once, whatever the sizes.  Host arithmetic only: the table comes from the
library (b200fdtd_split_geometry), the block -> cell mapping of locate_rect() is replayed here."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from mpifdtd_b200 import binding as B

K_BLOCK = 256


def cover(rects, shape):
    """Cells visited by one launch over `rects` (block numbering from 0), as a hit count."""
    hits = np.zeros(shape, dtype=np.int32)
    start = 0
    for r_lo, r_hi, c_lo, c_hi, lg, nbx, blk_end in rects:
        assert 4 <= lg <= 8 and blk_end > start
        rows_per_blk = K_BLOCK >> lg
        for b in range(blk_end - start):
            rb, cb = divmod(b, nbx)
            r0, c0 = r_lo + rb * rows_per_blk, c_lo + (cb << lg)
            rr = np.arange(r0, min(r0 + rows_per_blk, r_hi + 1))
            cc = np.arange(c0, min(c0 + (1 << lg), c_hi + 1))
            # every block must own at least one cell (no empty blocks beyond the ragged edge rule)
            assert len(rr) > 0 and len(cc) > 0
            hits[np.ix_(rr, cc)] += 1
        start = blk_end
    return hits


def geometry(updated, interior):
    L = B.lib()
    u = np.array(updated, dtype=np.int32)
    i = np.array(interior, dtype=np.int32)
    out = np.zeros(35, dtype=np.int32)
    n = B.C.c_int32(0)
    assert L.b200fdtd_split_geometry(u.ctypes.data, i.ctypes.data, out.ctypes.data, B.C.byref(n)) == 0
    return [tuple(int(v) for v in out[7 * q:7 * q + 7]) for q in range(n.value)]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(1, 40), st.integers(1, 700), st.integers(1, 700),
       st.integers(0, 30), st.integers(0, 30), st.integers(0, 300), st.integers(0, 300))
def test_interior_plus_frame_tile_the_updated_cells(plugin_lib, r_lo, c_lo, n_r, n_c, top, bottom, left, right):
    r_hi, c_hi = r_lo + n_r - 1, c_lo + n_c - 1
    ir_lo, ir_hi = min(r_lo + top, r_hi), max(r_hi - bottom, min(r_lo + top, r_hi))
    ic_lo, ic_hi = min(c_lo + left, c_hi), max(c_hi - right, min(c_lo + left, c_hi))
    rects = geometry((r_lo, r_hi, c_lo, c_hi), (ir_lo, ir_hi, ic_lo, ic_hi))
    assert rects[0][:4] == (ir_lo, ir_hi, ic_lo, ic_hi)
    shape = (r_hi + 2, c_hi + 2)
    hits = cover(rects[:1], shape) + cover(rects[1:], shape)
    want = np.zeros(shape, dtype=np.int32)
    want[r_lo:r_hi + 1, c_lo:c_hi + 1] = 1
    assert np.array_equal(hits, want)
    assert len(rects) <= 5


def test_benchmark_geometry(plugin_lib):
    """16384^2, pml 10: the interior is one rectangle of 256-wide blocks, the side strips get
    16 x 16-cell blocks instead of 245 idle threads per block."""
    n = 16384
    rects = geometry((2, n - 1, 9, n + 6), (11, n - 10, 18, n - 3))
    assert len(rects) == 5
    assert rects[0][4] == 8 and rects[0][6] == (n - 20) * ((n - 20 + 255) // 256)
    top, bottom, left, right = rects[1:]
    assert top[4] == 8 and bottom[4] == 8 and left[4] == 4 and right[4] == 4
    assert right[6] < 4000          # the whole frame in a few thousand blocks


def test_bad_rectangles_are_refused(plugin_lib):
    L = plugin_lib
    out = np.zeros(35, dtype=np.int32)
    n = B.C.c_int32(0)
    u = np.array([1, 10, 1, 10], dtype=np.int32)
    bad = np.array([0, 5, 1, 5], dtype=np.int32)
    assert L.b200fdtd_split_geometry(u.ctypes.data, bad.ctypes.data, out.ctypes.data, B.C.byref(n)) != 0

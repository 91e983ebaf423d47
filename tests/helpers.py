import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances (max-abs-diff / max-abs-ref), from BASELINE.json north_star / SURVEY 8(c):
TOL_FIELD = 1e-12      # field snapshots, double precision
TOL_FARFIELD = 1e-10   # U/W arrays and the 321 x 360 far-field table


def rel_err(a, b):
    """max |a-b| / max |b| (absolute if the reference is identically zero)."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    diff = np.abs(a - b).max()
    return diff / scale if scale > 0 else diff


def bit_equal(a, b):
    """Same shape and the same bit patterns (complex arrays: both components)."""
    a, b = np.asarray(a), np.asarray(b)
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        a = np.ascontiguousarray(a, dtype=np.complex128).view(np.float64)
        b = np.ascontiguousarray(b, dtype=np.complex128).view(np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


ANGLE_ROWS = list(range(0, 360, 24))
LAMBDA_ROWS = list(range(0, 321, 8))

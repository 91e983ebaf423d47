"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/liboracle.so, the plain-C
restatement of the reference's serial UPML path (oracle/fdtd_oracle.c) and of its four
split-field solvers (oracle/split_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
checker; never by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
TM, TE = 2, 3
TM_SLOTS = dict(Ez=0, Jz=1, Dz=2, Hx=3, Mx=4, Bx=5, Hy=6, My=7, By=8)
TE_SLOTS = dict(Ex=0, Jx=1, Dx=2, Ey=3, Jy=4, Dy=5, Hz=6, Mz=7, Bz=8)
TM_COEFS = ["C_JZ", "C_JZHXHY", "C_DZ", "C_DZJZ1", "C_DZJZ0", "C_MX", "C_MXEZ", "C_BX", "C_BXMX1",
            "C_BXMX0", "C_MY", "C_MYEZ", "C_BY", "C_BYMY1", "C_BYMY0"]
TE_COEFS = ["C_JX", "C_JXHZ", "C_DX", "C_DXJX1", "C_DXJX0", "C_JY", "C_JYHZ", "C_DY", "C_DYJY1",
            "C_DYJY0", "C_MZ", "C_MZEXEY", "C_BZ", "C_BZMZ1", "C_BZMZ0"]

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, "fdtd_oracle.c"), os.path.join(HERE, "split_oracle.c")]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in srcs):
            build()
        L = C.CDLL(LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_step_mpi.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_source_form.argtypes = [C.c_void_p, C.c_int]
        L.oracle_field.restype = C.c_void_p
        L.oracle_field.argtypes = [C.c_void_p, C.c_int]
        L.oracle_coef.restype = C.c_void_p
        L.oracle_coef.argtypes = [C.c_void_p, C.c_int]
        L.oracle_uw.restype = C.c_void_p
        L.oracle_uw.argtypes = [C.c_void_p, C.c_int]
        L.oracle_array_size.argtypes = [C.c_void_p]
        L.oracle_far_field.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_point_source.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_angle.argtypes = [C.c_void_p, C.c_int]
        L.oracle_ntff_box.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_fft.argtypes = [C.c_void_p, C.c_int]
        L.oracle_frequency_tm.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_time.restype = C.c_double
        L.oracle_time.argtypes = [C.c_void_p]
        L.split_oracle_create.restype = C.c_void_p
        L.split_oracle_create.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double] + [C.c_void_p] * 3
        L.split_oracle_destroy.argtypes = [C.c_void_p]
        L.split_oracle_step.argtypes = [C.c_void_p, C.c_int]
        L.split_oracle_field.restype = C.c_void_p
        L.split_oracle_field.argtypes = [C.c_void_p, C.c_int]
        L.split_oracle_coef.restype = C.c_void_p
        L.split_oracle_coef.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class OracleSim:
    def __init__(self, kind, n_px, n_py, steps, eps0, eps1=None, h_u_nm=10, pml=10, lambda_nm=500,
                 angle_deg=0, point_source=False, source_form=0):
        self.L = lib()
        self.kind, self.n_px, self.n_py, self.steps = kind, n_px, n_py, steps
        e0 = np.ascontiguousarray(eps0, dtype=np.float64)
        e1 = None if eps1 is None else np.ascontiguousarray(eps1, dtype=np.float64)
        assert e0.shape == (n_px, n_py)
        self.h = self.L.oracle_create(kind, n_px * h_u_nm, n_py * h_u_nm, h_u_nm, pml, lambda_nm,
                                      angle_deg, steps, e0.ctypes.data,
                                      None if e1 is None else e1.ctypes.data)
        if point_source:
            self.L.oracle_set_point_source(self.h, 1)
        if source_form:       # 1 "CW": field_scatteredWave instead of the pulse; 2 "PLANE": planeWave added
            self.L.oracle_set_source_form(self.h, {"CW": 1, "PLANE": 2}.get(source_form, source_form))
        self.array_size = self.L.oracle_array_size(self.h)

    def step(self, n=1, with_ntff=True):
        self.L.oracle_step(self.h, n, 1 if with_ntff else 0)

    def step_mpi(self, n=1, with_ntff=True):
        """n update() calls of the MPI-variant solver of this kind (id 4 for TM, 5 for TE) at one
        rank: E first, CW source, all cells against a zero ghost ring (mpiTM_UPML.c:196-217), then
        the variant's own ntff() into uw()."""
        self.L.oracle_step_mpi(self.h, n, 1 if with_ntff else 0)

    def field(self, name_or_slot):
        slots = TM_SLOTS if self.kind == TM else TE_SLOTS
        slot = slots[name_or_slot] if isinstance(name_or_slot, str) else name_or_slot
        n = self.n_px * self.n_py
        buf = (C.c_double * (2 * n)).from_address(self.L.oracle_field(self.h, slot))
        return np.frombuffer(buf, dtype=np.complex128).reshape(self.n_px, self.n_py).copy()

    def coef(self, name):
        names = TM_COEFS if self.kind == TM else TE_COEFS
        n = self.n_px * self.n_py
        buf = (C.c_double * n).from_address(self.L.oracle_coef(self.h, names.index(name)))
        return np.frombuffer(buf, dtype=np.float64).reshape(self.n_px, self.n_py).copy()

    def uw(self, slot):
        n = 360 * self.array_size
        buf = (C.c_double * (2 * n)).from_address(self.L.oracle_uw(self.h, slot))
        return np.frombuffer(buf, dtype=np.complex128).reshape(360, self.array_size).copy()

    def far_field(self):
        out = np.zeros((321, 360))
        self.L.oracle_far_field(self.h, out.ctypes.data)
        return out

    def frequency_tm(self):
        out = np.zeros(360, dtype=np.complex128)
        self.L.oracle_frequency_tm(self.h, out.ctypes.data)
        return out

    def box(self):
        b = (C.c_int * 6)()
        self.L.oracle_ntff_box(self.h, b)
        return dict(top=b[0], bottom=b[1], left=b[2], right=b[3], cx=b[4], cy=b[5])

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None


def fft(values):
    a = np.ascontiguousarray(values, dtype=np.complex128).copy()
    lib().oracle_fft(a.ctypes.data, a.size)
    return a


SPLIT_FIELDS = {0: ["Ez", "Ezx", "Ezy", "Hx", "Hy"], 6: ["Ez", "Ezx", "Ezy", "Hx", "Hy"],
                1: ["Hz", "Hzx", "Hzy", "Ex", "Ey"], 7: ["Hz", "Hzx", "Hzy", "Ex", "Ey"]}
SPLIT_COEFS = {0: ["C_EZX", "C_EZXLX", "C_EZY", "C_EZYLY", "C_HX", "C_HXLY", "C_HY", "C_HYLX"],
               1: ["C_EX", "C_EXLY", "C_EY", "C_EYLX", "C_HZX", "C_HZXLX", "C_HZY", "C_HZYLY"]}
SPLIT_COEFS[6], SPLIT_COEFS[7] = SPLIT_COEFS[0], SPLIT_COEFS[1]


class SplitOracleSim:
    """oracle/split_oracle.c: solver ids 0 (Yee TM + Berenger), 1 (TE), 6 (NS-FDTD TM), 7 (NS TE).
    eps_maps: the solver's three permittivity maps in the reference's order (TM kinds EPS_EZ,
    EPS_HX, EPS_HY; TE kinds EPS_EX, EPS_EY, EPS_HZ), each [n_px, n_py]."""

    def __init__(self, kind, n_px, n_py, eps_maps, h_u_nm=10, pml=10, lambda_nm=500, angle_deg=0):
        self.L = lib()
        self.kind, self.n_px, self.n_py = kind, n_px, n_py
        maps = [np.ascontiguousarray(m, dtype=np.float64) for m in eps_maps]
        assert len(maps) == 3 and all(m.shape == (n_px, n_py) for m in maps)
        # field_toCellUnit: nm / h_u in double (field.c:76)
        self.h = self.L.split_oracle_create(kind, n_px, n_py, pml, lambda_nm / float(h_u_nm), float(angle_deg),
                                            maps[0].ctypes.data, maps[1].ctypes.data, maps[2].ctypes.data)

    def step(self, n=1):
        self.L.split_oracle_step(self.h, n)

    def field(self, name):
        n = self.n_px * self.n_py
        buf = (C.c_double * (2 * n)).from_address(self.L.split_oracle_field(self.h, SPLIT_FIELDS[self.kind].index(name)))
        return np.frombuffer(buf, dtype=np.complex128).reshape(self.n_px, self.n_py).copy()

    def coef(self, name):
        n = self.n_px * self.n_py
        buf = (C.c_double * n).from_address(self.L.split_oracle_coef(self.h, SPLIT_COEFS[self.kind].index(name)))
        return np.frombuffer(buf, dtype=np.float64).reshape(self.n_px, self.n_py).copy()

    def close(self):
        if self.h:
            self.L.split_oracle_destroy(self.h)
            self.h = None

"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/_ref/libref.so.

libref.so is the UNMODIFIED reference (rennone/mpiFDTD) compiled from its own
sources by oracle/refbuild/Makefile.  It is the ground truth the parity tests
and the golden fixtures are pinned to, and the "reference" CPU baseline bench.py
times.  Nothing in the product (mpifdtd_b200/) imports this module.

The reference keeps all state in process globals (simulator.c:19-27,
field.c:11-34), so one process drives one simulation at a time; re-init after
finish is allowed (main.c:198-205).  The call order mirrors main.c:150-213:
models_setModel -> simulator_setSolver -> simulator_init(FieldInfo) ->
simulator_calc()* -> simulator_finish().
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBREF = os.path.join(HERE, "_ref", "libref.so")

# enum MODEL (models.h:5-14) and enum SOLVER (simulator.h:8-18)
MODELS = dict(NO_MODEL=0, MIE_CYLINDER=1, LAYER=2, MORPHO_SCALE=3,
              CONCENTRIC_CIRCLE=4, ZIGZAG=5, TRACE_IMAGE=6)
SOLVERS = dict(TM_2D=0, TE_2D=1, TM_UPML_2D=2, TE_UPML_2D=3,
               MPI_TM_UPML_2D=4, MPI_TE_UPML_2D=5, NS_TM_2D=6, NS_TE_2D=7)
HOOKS = {0: "refhook_tm", 1: "refhook_te", 2: "refhook_tm_upml", 3: "refhook_te_upml",
         4: "refhook_mpi_tm_upml", 5: "refhook_mpi_te_upml", 6: "refhook_ns_tm",
         7: "refhook_ns_te"}
MODE_D_X, MODE_D_Y, MODE_D_XY = 0, 1, 2


class FieldInfo(C.Structure):  # field.h:20-28
    _fields_ = [("width_nm", C.c_int), ("height_nm", C.c_int), ("h_u_nm", C.c_int),
                ("pml", C.c_int), ("lambda_nm", C.c_int), ("angle_deg", C.c_int),
                ("stepNum", C.c_int)]


class NTFFInfo(C.Structure):  # field.h:62-68
    _fields_ = [("top", C.c_int), ("bottom", C.c_int), ("left", C.c_int), ("right", C.c_int),
                ("cx", C.c_int), ("cy", C.c_int), ("RFperC", C.c_double),
                ("arraySize", C.c_int)]


def available():
    return os.path.exists(LIBREF)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so missing: run `make -C oracle/refbuild` "
                               "in the container that has /root/reference")
        L = C.CDLL(LIBREF)
        L.simulator_init.argtypes = [FieldInfo]
        L.models_setModel.argtypes = [C.c_int]
        L.simulator_setSolver.argtypes = [C.c_int]
        L.simulator_isFinish.restype = C.c_int
        L.models_eps.argtypes = [C.c_double, C.c_double, C.c_int]
        L.models_eps.restype = C.c_double
        L.field_sigmaX.argtypes = [C.c_double, C.c_double]
        L.field_sigmaX.restype = C.c_double
        L.field_sigmaY.argtypes = [C.c_double, C.c_double]
        L.field_sigmaY.restype = C.c_double
        L.field_getNTFFInfo.restype = NTFFInfo
        L.field_getTime.restype = C.c_double
        L.field_setWaveAngle.argtypes = [C.c_int]
        L.field_init.argtypes = [FieldInfo]
        for h in HOOKS.values():
            getattr(L, h).argtypes = [C.c_char_p]
            getattr(L, h).restype = C.c_void_p
        L.cfft.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class RefSim:
    """One reference simulation.  Output files land in `workdir` (a temp dir by
    default), because the reference writes its far-field files into cwd."""

    def __init__(self, model, solver, n_px, n_py=None, steps=100, h_u_nm=10, pml=10,
                 lambda_nm=500, angle_deg=0, workdir=None):
        self.L = lib()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        self.solver = SOLVERS[solver] if isinstance(solver, str) else int(solver)
        n_py = n_px if n_py is None else n_py
        self.n_px, self.n_py, self.steps = n_px, n_py, steps
        self.info = FieldInfo(n_px * h_u_nm, n_py * h_u_nm, h_u_nm, pml, lambda_nm,
                              angle_deg, steps)
        self.own_dir = workdir is None
        self.workdir = tempfile.mkdtemp(prefix="refsim_") if workdir is None else workdir
        self.prev_cwd = os.getcwd()
        os.chdir(self.workdir)
        os.makedirs("MPI_TM_UPML", exist_ok=True)   # mpiTM_UPML.c:175 writes there
        os.makedirs("MPI_TE_UPML", exist_ok=True)
        if self.model == MODELS["TRACE_IMAGE"]:
            # traceImageModel.c:8,82 opens "traceImage1.txt"; the shipped file is traceImage.txt
            src = os.environ.get("REFLIB_TRACE_IMAGE", "/root/reference/traceImage.txt")
            if os.path.exists(src) and not os.path.exists("traceImage1.txt"):
                shutil.copy(src, "traceImage1.txt")
        self.L.models_setModel(self.model)
        self.L.simulator_setSolver(self.solver)
        self.L.simulator_init(self.info)
        self.n_cell = n_px * n_py
        self.finished = False

    # -- stepping ---------------------------------------------------------
    def step(self, n=1):
        for _ in range(n):
            self.L.simulator_calc()

    def step_fn(self, fn_name, n=1):
        """Step with a refhook_* update variant followed by field_nextStep."""
        fn = getattr(self.L, fn_name)
        for _ in range(n):
            fn()
            self.L.field_nextStep()

    def run(self):
        while not self.L.simulator_isFinish():
            self.L.simulator_calc()

    def time(self):
        return self.L.field_getTime()

    def ntff_info(self):
        return self.L.field_getNTFFInfo()

    # -- array access -----------------------------------------------------
    def _ptr(self, name):
        p = getattr(self.L, HOOKS[self.solver])(name.encode())
        if not p:
            raise KeyError(name)
        return p

    def carray(self, name, count=None):
        """Copy of a complex file-static array (default N_CELL long), 1-D."""
        count = self.n_cell if count is None else count
        buf = (C.c_double * (2 * count)).from_address(self._ptr(name))
        return np.frombuffer(buf, dtype=np.complex128, count=count).copy()

    def darray(self, name, count=None):
        count = self.n_cell if count is None else count
        buf = (C.c_double * count).from_address(self._ptr(name))
        return np.frombuffer(buf, dtype=np.float64, count=count).copy()

    def field(self, name):
        """Complex field as [N_PX, N_PY] (index k = i*N_PY + j, field.c:70-72)."""
        return self.carray(name).reshape(self.n_px, self.n_py)

    def coef(self, name):
        return self.darray(name).reshape(self.n_px, self.n_py)

    def ntff_uw(self, name):
        a = self.ntff_info().arraySize
        return self.carray(name, 360 * a).reshape(360, a)

    # -- teardown ---------------------------------------------------------
    def finish(self):
        """simulator_finish(): writes "<ang>[deg]_380nm_700nm_b.dat" (UPML solvers)
        and frees.  Returns the 321x360 far-field table if one was written."""
        if self.finished:
            return None
        self.L.simulator_finish()
        self.finished = True
        out = None
        fn = os.path.join(self.workdir, "%d[deg]_380nm_700nm_b.dat" % self.info.angle_deg)
        if os.path.exists(fn):
            out = np.fromfile(fn, dtype=np.float64).reshape(321, 360)
        os.chdir(self.prev_cwd)
        if self.own_dir:
            shutil.rmtree(self.workdir, ignore_errors=True)
        return out


def eps_map(model, n_px, n_py, x_off, y_off, mode, h_u_nm=10, pml=10):
    """models_eps(i+x_off, j+y_off, mode) for every cell, via the reference's
    own dispatcher (models.c:132-150), without building a solver."""
    L = lib()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="refeps_")
    os.chdir(tmp)
    try:
        mid = MODELS[model] if isinstance(model, str) else int(model)
        if mid == MODELS["TRACE_IMAGE"]:
            shutil.copy(os.environ.get("REFLIB_TRACE_IMAGE", "/root/reference/traceImage.txt"),
                        "traceImage1.txt")
        L.models_setModel(mid)
        L.field_init(FieldInfo(n_px * h_u_nm, n_py * h_u_nm, h_u_nm, pml, 500, 0, 10))
        L.models_initModel()
        out = np.empty((n_px, n_py))
        for i in range(n_px):
            for j in range(n_py):
                out[i, j] = L.models_eps(i + x_off, j + y_off, mode)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return out

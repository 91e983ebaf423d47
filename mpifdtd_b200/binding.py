"""ctypes view of libmpifdtd_b200.so for the Python harness (tests, bench.py).

Python is plumbing here, not the product: the product is the C plugin surface
(include/mpifdtd_plugin.h) over the CUDA engine (include/b200fdtd.h).  This
module only declares argument types and wraps the two layers the way the
reference's own driver uses them (main.c:150-213):

    Plugin  -- models_setModel / simulator_setSolver / simulator_init / _calc /
               _reset / _finish and the borrowed-pointer getters
    Engine  -- the engine C ABI, for slab (multi-GPU) runs and kernel-level tests

There is no fallback: if the shared library is missing the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpifdtd_b200.so")

MODELS = dict(NO_MODEL=0, MIE_CYLINDER=1, LAYER=2, MORPHO_SCALE=3,
              CONCENTRIC_CIRCLE=4, ZIGZAG=5, TRACE_IMAGE=6)
SOLVERS = dict(TM_2D=0, TE_2D=1, TM_UPML_2D=2, TE_UPML_2D=3,
               MPI_TM_UPML_2D=4, MPI_TE_UPML_2D=5, NS_TM_2D=6, NS_TE_2D=7)
D_X, D_Y, D_XY = 0, 1, 2
UPML_TABS = 6
OPT_FUSED, OPT_STORE_H, OPT_BAND_ROWS, OPT_FUSED_SHAPE, OPT_F32_PAIRS = 1, 2, 3, 4, 7
OPT_DERIVED_E = 10
OPT_LEAN_INTERIOR = 8
OPT_UNIT_SPLIT = 9


class FieldInfo(C.Structure):
    _fields_ = [("width_nm", C.c_int), ("height_nm", C.c_int), ("h_u_nm", C.c_int),
                ("pml", C.c_int), ("lambda_nm", C.c_int), ("angle_deg", C.c_int),
                ("stepNum", C.c_int)]


class FieldInfoS(C.Structure):
    _fields_ = [("N_X", C.c_int), ("N_Y", C.c_int), ("N_PX", C.c_int), ("N_PY", C.c_int),
                ("N_CELL", C.c_int), ("N_PML", C.c_int), ("DX", C.c_int), ("DY", C.c_int)]


class NTFFInfo(C.Structure):
    _fields_ = [("top", C.c_int), ("bottom", C.c_int), ("left", C.c_int), ("right", C.c_int),
                ("cx", C.c_int), ("cy", C.c_int), ("RFperC", C.c_double), ("arraySize", C.c_int)]


class Grid(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_px", C.c_int32), ("n_py", C.c_int32), ("n_pml", C.c_int32),
                ("j0", C.c_int32), ("nj", C.c_int32), ("i_lo", C.c_int32), ("i_hi", C.c_int32),
                ("j_lo", C.c_int32), ("j_hi", C.c_int32), ("device", C.c_int32),
                ("precision", C.c_int32), ("mu0", C.c_double),
                ("n_batch", C.c_int32), ("flags", C.c_int32)]


class Pulse(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("reserved", C.c_int32), ("gap_x", C.c_double),
                ("gap_y", C.c_double), ("dot", C.c_double), ("cos_per_c", C.c_double),
                ("sin_per_c", C.c_double), ("time_minus_t0", C.c_double), ("omega", C.c_double),
                ("beam_width", C.c_double)]


class PointSource(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("i", C.c_int32), ("j", C.c_int32), ("reserved", C.c_int32),
                ("re", C.c_double), ("im", C.c_double)]


class CwSource(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("two_term", C.c_int32), ("gap_x", C.c_double),
                ("gap_y", C.c_double), ("scale", C.c_double), ("ks_cos", C.c_double),
                ("ks_sin", C.c_double), ("phase_a", C.c_double), ("phase_b", C.c_double)]


class BatchCw(C.Structure):                         # b200fdtd_batch_cw
    _fields_ = [("ks_cos", C.c_double), ("ks_sin", C.c_double), ("dot", C.c_double * 2), ("enabled", C.c_int32 * 2)]


class LineSource(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("i", C.c_int32), ("j_lo", C.c_int32), ("j_hi", C.c_int32),
                ("scale", C.c_double), ("ks_cos", C.c_double), ("ks_sin", C.c_double),
                ("time", C.c_double), ("omega", C.c_double)]


class StepArgs(C.Structure):
    _fields_ = [("time", C.c_double), ("ray_coef", C.c_double), ("pulse", Pulse * 2),
                ("point", PointSource), ("cw", CwSource * 2), ("ns_r2", C.c_double),
                ("line", LineSource)]


class NtffPlan(C.Structure):
    _fields_ = [("top", C.c_int32), ("bottom", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("n_points", C.c_int32), ("n_local", C.c_int32), ("max_time", C.c_int32),
                ("n_bins", C.c_int32), ("n_angles", C.c_int32), ("array_size", C.c_int32),
                ("time_shift", C.c_void_p), ("tap_scale", C.c_double),
                ("sample_di", C.c_int32), ("sample_dj", C.c_int32)]


class SpectrumArgs(C.Structure):
    _fields_ = [("coef_re", C.c_double), ("coef_im", C.c_double), ("z0", C.c_double),
                ("cos_phi", C.c_void_p), ("sin_phi", C.c_void_p), ("n_fft", C.c_int32),
                ("lambda_first_nm", C.c_int32), ("lambda_last_nm", C.c_int32),
                ("reserved", C.c_int32), ("c_hu_nfft", C.c_double), ("twiddle", C.c_void_p)]


C_0_S = 0.7071
MU_0_S = 1.0 / C_0_S / C_0_S
Z_0_S = 1.41422712488

_lib = None


def lib():
    """The shared library, loaded once.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.b200fdtd_last_error.restype = C.c_char_p
    L.b200fdtd_create.argtypes = [C.POINTER(Grid), C.POINTER(vp)]
    L.b200fdtd_destroy.argtypes = [vp]
    L.b200fdtd_device_count.argtypes = [C.POINTER(C.c_int)]
    L.b200fdtd_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.b200fdtd_host_free.argtypes = [vp]
    L.b200fdtd_set_upml_tables.argtypes = [vp, vp, vp]
    L.b200fdtd_set_eps.argtypes = [vp, i32, vp]
    L.b200fdtd_set_ntff_plan.argtypes = [vp, C.POINTER(NtffPlan)]
    for fn in ("b200fdtd_step", "b200fdtd_phase_h", "b200fdtd_phase_e", "b200fdtd_phase_sample",
               "b200fdtd_phase_fused"):
        getattr(L, fn).argtypes = [vp, C.POINTER(StepArgs)]
    L.b200fdtd_sync.argtypes = [vp]
    L.b200fdtd_halo_pack.argtypes = [vp, i32, vp]
    L.b200fdtd_halo_unpack.argtypes = [vp, i32, vp]
    L.b200fdtd_set_stream.argtypes = [vp, vp]
    L.b200fdtd_get_field.argtypes = [vp, i32, vp]
    L.b200fdtd_set_field.argtypes = [vp, i32, vp]
    L.b200fdtd_get_field_slab.argtypes = [vp, i32, vp]
    L.b200fdtd_set_option.argtypes = [vp, i32, i32]
    L.b200fdtd_set_dense.argtypes = [vp, i32, vp]
    L.b200fdtd_set_split_interior.argtypes = [vp, i32, i32, i32, i32]
    L.b200fdtd_peer_export.argtypes = [vp, vp]
    L.b200fdtd_peer_attach.argtypes = [vp, i32, vp]
    L.b200fdtd_peer_attach_engine.argtypes = [vp, i32, vp]
    L.b200fdtd_ntff_add_uw.argtypes = [vp, vp]
    L.mpifdtd_ntffFrequency.argtypes = [C.c_int, vp]
    L.mpifdtd_split_prepare_host.argtypes = [C.c_int]
    L.mpifdtd_split_prepare_host_lean.argtypes = [C.c_int]
    L.mpifdtd_split_lean_tables.argtypes = [C.c_int, vp, vp]
    L.mpifdtd_split_dense.argtypes = [C.c_int, C.c_int]
    L.mpifdtd_split_dense.restype = vp
    L.mpifdtd_split_engine.argtypes = [C.c_int]
    L.mpifdtd_split_engine.restype = vp
    L.mpifdtd_split_step_args.argtypes = [C.c_int, C.POINTER(StepArgs)]
    L.b200fdtd_selftest_division.argtypes = [dbl, C.c_uint64, C.POINTER(C.c_uint64)]
    L.b200fdtd_zero_state.argtypes = [vp]
    L.b200fdtd_field_digest.argtypes = [vp, i32, C.POINTER(C.c_uint64)]
    L.b200fdtd_ntff_project.argtypes = [vp]
    L.b200fdtd_ntff_get_uw.argtypes = [vp, i32, vp]
    L.b200fdtd_ntff_uw_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.b200fdtd_ntff_spectrum.argtypes = [vp, C.POINTER(SpectrumArgs), vp]
    L.b200fdtd_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200fdtd_onepass_vacuum_cells.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200fdtd_device_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.b200fdtd_timer_start.argtypes = [vp]
    L.b200fdtd_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    # plugin surface
    L.field_init.argtypes = [FieldInfo]
    L.simulator_init.argtypes = [FieldInfo]
    L.models_setModel.argtypes = [C.c_int]
    L.simulator_setSolver.argtypes = [C.c_int]
    L.simulator_isFinish.restype = C.c_int
    L.simulator_getEps.restype = vp
    L.simulator_getDrawingData.restype = vp
    L.models_eps.argtypes = [dbl, dbl, C.c_int]
    L.models_eps.restype = dbl
    L.field_sigmaX.argtypes = [dbl, dbl]
    L.field_sigmaX.restype = dbl
    L.field_sigmaY.argtypes = [dbl, dbl]
    L.field_sigmaY.restype = dbl
    L.field_getNTFFInfo.restype = NTFFInfo
    L.field_getFieldInfo_S.restype = FieldInfoS
    for fn in ("field_getTime", "field_getMaxTime", "field_getOmega", "field_getK",
               "field_getRayCoef", "field_getWaveAngle", "field_getLambda", "field_getT"):
        getattr(L, fn).restype = dbl
    L.field_setWaveAngle.argtypes = [C.c_int]
    L.mpifdtd_fill_eps.argtypes = [vp, dbl, dbl, C.c_int]
    L.mpifdtd_upml_dense_coefficient.argtypes = [C.c_int, C.c_char_p, vp]
    L.mpifdtd_ntff_time_shift.argtypes = [C.POINTER(NTFFInfo), C.c_int, dbl, C.c_int, C.c_int]
    L.mpifdtd_ntff_local_count.argtypes = [C.POINTER(NTFFInfo), C.c_int, C.c_int]
    L.mpifdtd_fill_eps_slab.argtypes = [vp, dbl, dbl, C.c_int, C.c_int, C.c_int]
    L.mpifdtd_upml_tables.argtypes = [C.c_int, vp, vp]
    L.b200fdtd_set_eps_slab.argtypes = [vp, i32, vp]
    L.b200fdtd_set_eps_palette.argtypes = [vp, i32, vp, C.c_int64, vp, i32]
    L.mpifdtd_eps_palette.argtypes = [vp, C.c_size_t, vp, vp]
    L.mpifdtd_upml_step_args.argtypes = [C.c_int, C.c_int, C.POINTER(StepArgs)]
    L.mpifdtd_upml_far_field.argtypes = [vp, C.c_int, C.c_int, vp]
    L.mpifdtd_ntff_time_shift.restype = vp
    L.mpifdtd_ntff_point_count.argtypes = [C.POINTER(NTFFInfo)]
    L.mpifdtd_fft_twiddles.argtypes = [C.c_int]
    L.mpifdtd_fft_twiddles.restype = vp
    L.mpifdtd_ntff_direction_cosines.argtypes = [C.c_int, C.c_int, vp, vp]
    L.mpifdtd_upml_engine.argtypes = [C.c_int]
    L.mpifdtd_upml_engine.restype = vp
    L.mpifdtd_enablePointSource.argtypes = [C.c_int]
    L.mpifdtd_setSourceForm.argtypes = [C.c_int]
    L.mpifdtd_setPrecision.argtypes = [C.c_int]
    L.mpifdtd_setAngleBatch.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.mpifdtd_selectAngle.argtypes = [C.c_int]
    L.mpifdtd_runAngleSweep.argtypes = [FieldInfo, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200fdtd_mem_info.argtypes = [i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.b200fdtd_select_batch.argtypes = [vp, i32]
    L.b200fdtd_run_steps.argtypes = [vp, dbl, i32]
    L.b200fdtd_set_batch_sources.argtypes = [vp, vp]
    L.b200fdtd_struct_size.argtypes = [i32]
    L.b200fdtd_upml_interior.argtypes = [i32, vp, i32, vp, i32, vp]
    L.b200fdtd_get_lean_extent.argtypes = [vp, vp]
    L.b200fdtd_get_step_form.argtypes = [vp, C.POINTER(i32)]
    L.b200fdtd_split_geometry.argtypes = [vp, vp, vp, C.POINTER(i32)]
    L.mpifdtd_readConfig.argtypes = [C.c_char_p, vp]
    for name in ("fdtdTM_upml_getHx", "fdtdTM_upml_getHy", "fdtdTM_upml_getEz",
                 "fdtdTE_upml_getEx", "fdtdTE_upml_getEy", "fdtdTE_upml_getHz",
                 "fdtdTM_upml_getEps", "fdtdTE_upml_getEps"):
        getattr(L, name).restype = vp
    for prefix, fields in (("mpi_fdtdTM_upml", ("Hx", "Hy", "Ez")), ("mpi_fdtdTE_upml", ("Ex", "Ey", "Hz")),
                           ("fdtdTM", ("Hx", "Hy", "Ez", "Ezx", "Ezy")), ("fdtdTE", ("Ex", "Ey", "Hz", "Hzx", "Hzy")),
                           ("nsFdtdTM", ("Hx", "Hy", "Ez", "Ezx", "Ezy")),
                           ("nsFdtdTE", ("Ex", "Ey", "Hz", "Hzx", "Hzy"))):
        for f in fields + ("Eps",):
            getattr(L, "%s_get%s" % (prefix, f)).restype = vp
    L.free.argtypes = [vp]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise EngineError("%s failed (%d): %s" % (what, rc, lib().b200fdtd_last_error().decode()))


def device_count():
    n = C.c_int(0)
    lib().b200fdtd_device_count(C.byref(n))
    return n.value


def _as_complex(ptr, n_px, n_py):
    buf = (C.c_double * (2 * n_px * n_py)).from_address(ptr)
    return np.frombuffer(buf, dtype=np.complex128).reshape(n_px, n_py)


PRECISIONS = dict(f64=0, f32=1)                     # B200FDTD_F64 / B200FDTD_F32
SOURCE_FORMS = dict(DEFAULT=0, CW=1, PLANE=2)       # MPIFDTD_SRC_* of mpifdtd_plugin.h


class Plugin:
    """The reference's driver sequence against the plugin surface."""

    GETTERS = {2: dict(Hx="fdtdTM_upml_getHx", Hy="fdtdTM_upml_getHy", Ez="fdtdTM_upml_getEz"),
               3: dict(Ex="fdtdTE_upml_getEx", Ey="fdtdTE_upml_getEy", Hz="fdtdTE_upml_getHz"),
               4: {f: "mpi_fdtdTM_upml_get" + f for f in ("Hx", "Hy", "Ez")},
               5: {f: "mpi_fdtdTE_upml_get" + f for f in ("Ex", "Ey", "Hz")},
               0: {f: "fdtdTM_get" + f for f in ("Hx", "Hy", "Ez", "Ezx", "Ezy")},
               1: {f: "fdtdTE_get" + f for f in ("Ex", "Ey", "Hz", "Hzx", "Hzy")},
               6: {f: "nsFdtdTM_get" + f for f in ("Hx", "Hy", "Ez", "Ezx", "Ezy")},
               7: {f: "nsFdtdTE_get" + f for f in ("Ex", "Ey", "Hz", "Hzx", "Hzy")}}

    def __init__(self, model, solver, n_px, n_py=None, steps=100, h_u_nm=10, pml=10,
                 lambda_nm=500, angle_deg=0, point_source=False, source_form=0, precision=0, angle_batch=None):
        self.L = lib()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        self.solver = SOLVERS[solver] if isinstance(solver, str) else int(solver)
        n_py = n_px if n_py is None else n_py
        self.n_px, self.n_py, self.steps = n_px, n_py, steps
        self.info = FieldInfo(n_px * h_u_nm, n_py * h_u_nm, h_u_nm, pml, lambda_nm, angle_deg, steps)
        self.L.mpifdtd_enablePointSource(1 if point_source else 0)
        self.L.mpifdtd_setSourceForm(SOURCE_FORMS[source_form] if isinstance(source_form, str) else source_form)
        self.L.mpifdtd_setPrecision(PRECISIONS[precision] if isinstance(precision, str) else precision)
        self.angle_batch = list(angle_batch) if angle_batch else []
        if self.angle_batch:      # all these incidence angles at once, one batched engine
            arr = (C.c_int * len(self.angle_batch))(*self.angle_batch)
            self.L.mpifdtd_setAngleBatch(arr, len(self.angle_batch))
        else:
            self.L.mpifdtd_setAngleBatch(None, 0)
        self.L.models_setModel(self.model)
        self.L.simulator_setSolver(self.solver)
        self.L.simulator_init(self.info)
        self.finished = False

    def engine_handle(self):
        if self.solver in (2, 3, 4, 5):
            return self.L.mpifdtd_upml_engine(self.solver)
        return self.L.mpifdtd_split_engine(self.solver)

    def select_angle(self, index):
        """Which simulation of an angle batch the getters refer to."""
        self.L.mpifdtd_selectAngle(index)

    def far_field_file(self, angle_deg, workdir="."):
        fn = os.path.join(workdir, "%d[deg]_380nm_700nm_b.dat" % angle_deg)
        return np.fromfile(fn, dtype=np.float64).reshape(321, 360)

    def step(self, n=1):
        for _ in range(n):
            self.L.simulator_calc()

    def run(self):
        while not self.L.simulator_isFinish():
            self.L.simulator_calc()

    def sync(self):
        check(self.L.b200fdtd_sync(self.engine_handle()), "sync")

    SLOTS = {2: ["Ez", "Jz", "Dz", "Hx", "Mx", "Bx", "Hy", "My", "By"],
             3: ["Ex", "Jx", "Dx", "Ey", "Jy", "Dy", "Hz", "Mz", "Bz"]}
    SLOTS[4], SLOTS[5] = SLOTS[2], SLOTS[3]

    def field(self, name):
        """The three public fields come through the reference's getters (borrowed host
        mirror); the auxiliary J/D/M/B arrays, which the reference keeps file-static,
        through the engine's slot accessor.  The MPI-variant ids hand out (N+2) x (N+2)
        arrays with a ghost ring, like the reference."""
        if name in self.GETTERS[self.solver]:
            ptr = getattr(self.L, self.GETTERS[self.solver][name])()
            if self.solver in (4, 5):
                return _as_complex(ptr, self.n_px + 2, self.n_py + 2).copy()
            return _as_complex(ptr, self.n_px, self.n_py).copy()
        return self.any_field(self.SLOTS[self.solver].index(name))

    def any_field(self, slot):
        out = np.zeros((self.n_px, self.n_py), dtype=np.complex128)
        check(self.L.b200fdtd_get_field(self.engine_handle(), slot, out.ctypes.data), "get_field")
        return out

    def eps(self):
        ptr = self.L.simulator_getEps()
        buf = (C.c_double * (self.n_px * self.n_py)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64).reshape(self.n_px, self.n_py).copy()

    def ntff_uw(self, slot, project=True):
        h = self.engine_handle()
        if project:
            check(self.L.b200fdtd_ntff_project(h), "ntff_project")
        n_bins = self.ntff_bins()
        out = np.zeros((360, n_bins), dtype=np.complex128)
        check(self.L.b200fdtd_ntff_get_uw(h, slot, out.ctypes.data), "ntff_get_uw")
        return out

    def ntff_bins(self):
        if os.environ.get("MPIFDTD_NTFF_FULL_BINS"):
            return self.L.field_getNTFFInfo().arraySize
        return self.steps

    def launches(self):
        n = C.c_uint64(0)
        self.L.b200fdtd_launch_count(self.engine_handle(), C.byref(n))
        return n.value

    def lean_extent(self):
        """(i_lo, i_hi, j_lo, j_hi) of the cells B200FDTD_OPT_LEAN_INTERIOR treats as frame-free
        (inclusive; lo > hi when the option is off)."""
        out = (C.c_int32 * 4)()
        check(self.L.b200fdtd_get_lean_extent(self.engine_handle(), out), "get_lean_extent")
        return tuple(out)

    def finish(self, workdir=None):
        """simulator_finish(); returns the 321 x 360 table written to cwd."""
        if self.finished:
            return None
        cwd = os.getcwd()
        if workdir is not None:
            os.chdir(workdir)
        try:
            self.L.simulator_finish()
            fn = "%d[deg]_380nm_700nm_b.dat" % self.info.angle_deg
            out = np.fromfile(fn, dtype=np.float64).reshape(321, 360) if os.path.exists(fn) else None
        finally:
            os.chdir(cwd)
        self.finished = True
        return out


def upml_tables(kind, n_px, n_py):
    """1-D coefficient tables for the current field_init() state, rebuilt from the
    dense-coefficient probe (tests) -- production code builds them in upml_shim.c."""
    raise NotImplementedError


class Engine:
    """Engine-level access (one y-slab).  Host-side preparation (eps maps,
    coefficient tables, NTFF plan) comes from the plugin's own C helpers after
    field_init(), so a slab engine sees exactly what the serial shim would."""

    def __init__(self, kind, n_px, n_py, n_pml, j0=0, nj=None, device=-1, extents=None, precision=0):
        self.L = lib()
        precision = PRECISIONS[precision] if isinstance(precision, str) else precision
        nj = n_py - j0 if nj is None else nj
        i_lo, i_hi, j_lo, j_hi = extents if extents else (1, n_px - 2, 1, n_py - 2)
        self.grid = Grid(kind, n_px, n_py, n_pml, j0, nj, i_lo, i_hi, j_lo, j_hi, device, precision, MU_0_S)
        self.h = C.c_void_p()
        check(self.L.b200fdtd_create(C.byref(self.grid), C.byref(self.h)), "create")
        self.kind, self.n_px, self.n_py, self.j0, self.nj = kind, n_px, n_py, j0, nj
        self.n_bins = 0

    def close(self):
        if self.h:
            self.L.b200fdtd_destroy(self.h)
            self.h = C.c_void_p()

    def set_tables(self, tab_i, tab_j):
        ti = np.ascontiguousarray(tab_i, dtype=np.float64)
        tj = np.ascontiguousarray(tab_j, dtype=np.float64)
        assert ti.shape == (UPML_TABS, self.n_px) and tj.shape == (UPML_TABS, self.n_py)
        check(self.L.b200fdtd_set_upml_tables(self.h, ti.ctypes.data, tj.ctypes.data), "set_tables")

    def set_eps(self, slot, eps):
        e = np.ascontiguousarray(eps, dtype=np.float64)
        assert e.shape == (self.n_px, self.n_py)
        check(self.L.b200fdtd_set_eps(self.h, slot, e.ctypes.data), "set_eps")

    def set_ntff(self, box, time_shift, max_time, n_bins=None, n_angles=360, array_size=None):
        ts = np.ascontiguousarray(time_shift, dtype=np.float64)
        n_points = 2 * (box.right - box.left) + 2 * (box.top - box.bottom)
        assert ts.shape == (n_angles, n_points)
        self.n_bins = max_time if n_bins is None else n_bins
        array_size = box.arraySize if array_size is None else array_size
        plan = NtffPlan(top=box.top, bottom=box.bottom, left=box.left, right=box.right,
                        n_points=n_points, n_local=n_points, max_time=max_time, n_bins=self.n_bins,
                        n_angles=n_angles, array_size=array_size, time_shift=ts.ctypes.data)
        check(self.L.b200fdtd_set_ntff_plan(self.h, C.byref(plan)), "set_ntff_plan")

    def step(self, args):
        check(self.L.b200fdtd_step(self.h, C.byref(args)), "step")

    def phase_h(self, args):
        check(self.L.b200fdtd_phase_h(self.h, C.byref(args)), "phase_h")

    def phase_e(self, args):
        check(self.L.b200fdtd_phase_e(self.h, C.byref(args)), "phase_e")

    def phase_fused(self, args):
        check(self.L.b200fdtd_phase_fused(self.h, C.byref(args)), "phase_fused")

    def phase_sample(self, args):
        check(self.L.b200fdtd_phase_sample(self.h, C.byref(args)), "phase_sample")

    def sync(self):
        check(self.L.b200fdtd_sync(self.h), "sync")

    def set_stream(self, stream_handle):
        check(self.L.b200fdtd_set_stream(self.h, C.c_void_p(stream_handle)), "set_stream")

    def peer_export(self):
        blob = C.create_string_buffer(256)
        check(self.L.b200fdtd_peer_export(self.h, blob), "peer_export")
        return blob.raw

    def peer_attach(self, which_neighbour, blob):
        buf = C.create_string_buffer(blob, 256)
        check(self.L.b200fdtd_peer_attach(self.h, which_neighbour, buf), "peer_attach")

    def peer_attach_engine(self, which_neighbour, other):
        """Same-process attachment (several slabs / devices under one host thread)."""
        check(self.L.b200fdtd_peer_attach_engine(self.h, which_neighbour, other.h), "peer_attach_engine")

    def add_uw(self, other):
        check(self.L.b200fdtd_ntff_add_uw(self.h, other.h), "ntff_add_uw")

    def halo_pack(self, which, dev_ptr):
        check(self.L.b200fdtd_halo_pack(self.h, which, C.c_void_p(dev_ptr)), "halo_pack")

    def halo_unpack(self, which, dev_ptr):
        check(self.L.b200fdtd_halo_unpack(self.h, which, C.c_void_p(dev_ptr)), "halo_unpack")

    def get_field(self, slot, out=None):
        if out is None:
            out = np.zeros((self.n_px, self.n_py), dtype=np.complex128)
        check(self.L.b200fdtd_get_field(self.h, slot, out.ctypes.data), "get_field")
        return out

    def get_field_slab(self, slot):
        out = np.zeros((self.n_px, self.nj), dtype=np.complex128)
        check(self.L.b200fdtd_get_field_slab(self.h, slot, out.ctypes.data), "get_field_slab")
        return out

    def set_field(self, slot, values):
        v = np.ascontiguousarray(values, dtype=np.complex128)
        assert v.shape == (self.n_px, self.n_py)
        check(self.L.b200fdtd_set_field(self.h, slot, v.ctypes.data), "set_field")

    def zero(self):
        check(self.L.b200fdtd_zero_state(self.h), "zero_state")

    def digest(self, slot):
        """64-bit position-mixed digest of this slab's cells of a field (sums over slabs mod 2^64)."""
        d = C.c_uint64(0)
        check(self.L.b200fdtd_field_digest(self.h, slot, C.byref(d)), "field_digest")
        return d.value

    def set_option(self, option, value):
        check(self.L.b200fdtd_set_option(self.h, option, value), "set_option")

    def vacuum_cells(self):
        """cells per step in vacuum row-strips of the one-pass step (no E arrays kept there)"""
        n = C.c_uint64(0)
        check(self.L.b200fdtd_onepass_vacuum_cells(self.h, C.byref(n)), "onepass_vacuum_cells")
        return n.value

    def step_form(self):
        """0 one full kernel per phase, 1 unit-coefficient interior + frame, 2 lean interior + frame,
        3 the one-pass step (phase_h / phase_e then still launch form 0 or 1), 4 one pass, lean form."""
        form = C.c_int32(-1)
        check(self.L.b200fdtd_get_step_form(self.h, C.byref(form)), "get_step_form")
        return form.value

    def project(self):
        check(self.L.b200fdtd_ntff_project(self.h), "ntff_project")

    def uw(self, slot, n_angles=360):
        out = np.zeros((n_angles, self.n_bins), dtype=np.complex128)
        check(self.L.b200fdtd_ntff_get_uw(self.h, slot, out.ctypes.data), "ntff_get_uw")
        return out

    def uw_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(self.L.b200fdtd_ntff_uw_device(self.h, C.byref(p), C.byref(n)), "uw_device")
        return p.value, n.value

    def spectrum(self, args):
        n_lam = args.lambda_last_nm - args.lambda_first_nm + 1
        out = np.zeros((n_lam, 360))
        check(self.L.b200fdtd_ntff_spectrum(self.h, C.byref(args), out.ctypes.data), "ntff_spectrum")
        return out

    def launches(self):
        n = C.c_uint64(0)
        self.L.b200fdtd_launch_count(self.h, C.byref(n))
        return n.value

    def device_bytes(self):
        n = C.c_uint64(0)
        self.L.b200fdtd_device_bytes(self.h, C.byref(n))
        return n.value

    def timer_start(self):
        check(self.L.b200fdtd_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        check(self.L.b200fdtd_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

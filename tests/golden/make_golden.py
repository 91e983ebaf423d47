"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libref.so):

    make -C oracle/refbuild && python tests/golden/make_golden.py

Every array below is produced by the reference's own code (driven through
oracle/reflib.py); nothing here is computed by this repository's product or by
the C restatement.  The fixtures are what pins the oracle (tests/test_oracle_cpu.py)
and the product's host-side maps (tests/test_materials_cpu.py) on machines that do
not have /root/reference, e.g. the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reflib  # noqa: E402

# the trace-image model reads an index image; the fixture is our own synthetic one
os.environ["REFLIB_TRACE_IMAGE"] = os.path.join(HERE, "traceImage_fixture.txt")

ANGLE_ROWS = list(range(0, 360, 24))         # 15 directions of U/W kept
LAMBDA_ROWS = list(range(0, 321, 8))         # 41 wavelength rows of the far field kept


def eps_fixtures():
    """models_eps maps: every model, the sampling modes the UPML solvers use."""
    out = {}
    cases = [("NO_MODEL", 40, 44, 10), ("MIE_CYLINDER", 144, 150, 10), ("LAYER", 90, 140, 10),
             ("MORPHO_SCALE", 90, 190, 10), ("ZIGZAG", 90, 150, 10), ("TRACE_IMAGE", 96, 100, 10),
             ("MIE_CYLINDER", 100, 96, 20)]
    modes = [("ez", 0.0, 0.0, reflib.MODE_D_XY), ("ex", 0.5, 0.0, reflib.MODE_D_Y),
             ("ey", 0.0, 0.5, reflib.MODE_D_X), ("hzx", 0.5, 0.5, reflib.MODE_D_X)]
    for model, npx, npy, hu in cases:
        for tag, xo, yo, mode in modes:
            key = "%s_%dx%d_hu%d_%s" % (model, npx, npy, hu, tag)
            out[key] = reflib.eps_map(model, npx, npy, xo, yo, mode, h_u_nm=hu)
            print(key, int((out[key] != 1).sum()), "non-vacuum cells")
    np.savez_compressed(os.path.join(HERE, "eps_maps.npz"), **out)


def grid_fixtures():
    """field_init integers (grid, NTFF box), the PML profile and wave scalars."""
    L = reflib.lib()
    L.field_getOmega.restype = reflib.C.c_double
    L.field_getK.restype = reflib.C.c_double
    L.field_getRayCoef.restype = reflib.C.c_double
    rows = []
    for (w, h, hu, pml, lam, steps) in [(2560, 2560, 10, 10, 500, 2000), (10240, 10240, 10, 10, 500, 2000),
                                        (1930, 2570, 10, 15, 633, 700), (2000, 1900, 20, 10, 500, 640),
                                        (5005, 7777, 7, 12, 450, 100)]:
        L.field_init(reflib.FieldInfo(w, h, hu, pml, lam, 0, steps))
        b = L.field_getNTFFInfo()
        npx, npy = int(reflib.C.c_int.in_dll(L, "N_PX").value), int(reflib.C.c_int.in_dll(L, "N_PY").value)
        xs = [0, 0.5, 1, pml - 1, pml - 0.5, pml, pml + 0.5, npx // 2, npx - pml - 1.5, npx - pml - 1,
              npx - pml - 0.5, npx - pml, npx - 1.5, npx - 1, npx - 0.5]
        ys = [0, 0.5, pml - 0.5, pml, npy // 2, npy - pml - 1, npy - pml - 0.5, npy - pml, npy - 1, npy - 0.5]
        ramp = []
        L.field_reset()
        for _ in range(5):
            L.field_nextStep()
            ramp.append(L.field_getRayCoef())
        rows.append(dict(width_nm=w, height_nm=h, h_u_nm=hu, pml=pml, lambda_nm=lam, steps=steps,
                         N_PX=npx, N_PY=npy, top=b.top, bottom=b.bottom, left=b.left, right=b.right,
                         cx=b.cx, cy=b.cy, RFperC=b.RFperC, arraySize=b.arraySize,
                         omega=L.field_getOmega().hex(), k=L.field_getK().hex(),
                         sigma_x=[[x, L.field_sigmaX(x, 0.0).hex()] for x in xs],
                         sigma_y=[[y, L.field_sigmaY(0.0, y).hex()] for y in ys],
                         ray_coef=[r.hex() for r in ramp]))
    with open(os.path.join(HERE, "grid_cases.json"), "w") as f:
        json.dump(rows, f, indent=1)


def run_fixture(tag, model, solver, npx, npy, hu, steps, angle, fields, uw_names, eps_names, coef_names):
    sim = reflib.RefSim(model, solver, npx, npy, steps=steps, h_u_nm=hu, angle_deg=angle)
    out = {"meta": np.array([npx, npy, hu, steps, angle, reflib.SOLVERS[solver], reflib.MODELS[model]])}
    for e in eps_names:
        out[e] = sim.coef(e)
    for c in coef_names:
        out[c] = sim.coef(c)
    half = steps // 2
    sim.step(half)
    for f in fields:
        out["mid_" + f] = sim.field(f)
    sim.step(steps - half)
    for f in fields:
        out["end_" + f] = sim.field(f)
    for u in uw_names:
        out["uw_" + u] = sim.ntff_uw(u)[ANGLE_ROWS, :steps]
    ff = sim.finish()
    out["far_field_rows"] = ff[LAMBDA_ROWS, :]
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, "far field max", ff.max())


def run_fixtures():
    tm_f = ["Ez", "Hx", "Hy", "Jz", "Dz", "Mx", "Bx", "My", "By"]
    te_f = ["Ex", "Ey", "Hz", "Jx", "Dx", "Jy", "Dy", "Mz", "Bz"]
    tm_c = ["C_JZ", "C_JZHXHY", "C_DZ", "C_DZJZ1", "C_DZJZ0", "C_MX", "C_MXEZ", "C_BX", "C_BXMX1", "C_BXMX0",
            "C_MY", "C_MYEZ", "C_BY", "C_BYMY1", "C_BYMY0"]
    te_c = ["C_JX", "C_JXHZ", "C_DX", "C_DXJX1", "C_DXJX0", "C_JY", "C_JYHZ", "C_DY", "C_DYJY1", "C_DYJY0",
            "C_MZ", "C_MZEXEY", "C_BZ", "C_BZMZ1", "C_BZMZ0"]
    run_fixture("mie_tm_upml_88x96", "MIE_CYLINDER", "TM_UPML_2D", 88, 96, 20, 640, 0,
                tm_f[:3], ["Ux", "Uy", "Wz"], ["EPS_EZ"], tm_c)
    run_fixture("mie_te_upml_88x96", "MIE_CYLINDER", "TE_UPML_2D", 88, 96, 20, 640, 0,
                te_f[:1] + te_f[3:4] + te_f[6:7], ["Wx", "Wy", "Uz"], ["EPS_EX", "EPS_EY"], te_c)
    run_fixture("zigzag_tm_upml_72x120_a30", "ZIGZAG", "TM_UPML_2D", 72, 120, 10, 620, 30,
                tm_f[:3], ["Wz"], ["EPS_EZ"], [])
    run_fixture("layer_te_upml_80x110_a45", "LAYER", "TE_UPML_2D", 80, 110, 10, 620, 45,
                ["Ex", "Ey", "Hz"], ["Uz"], ["EPS_EX", "EPS_EY"], [])


def fft_fixture():
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(256) + 1j * rng.standard_normal(256)).astype(np.complex128)
    y = x.copy()
    reflib.lib().cfft(y.ctypes.data, 256)
    np.savez_compressed(os.path.join(HERE, "cfft_256.npz"), x=x, y=y)


if __name__ == "__main__":
    which = sys.argv[1:] or ["eps", "grid", "fft", "runs"]
    if "eps" in which:
        eps_fixtures()
    if "grid" in which:
        grid_fixtures()
    if "fft" in which:
        fft_fixture()
    if "runs" in which:
        run_fixtures()


def split_fixtures():
    """Split-field solvers (ids 0, 1, 6, 7): dense coefficients, eps maps and field
    snapshots of a Mie-cylinder run (CW source, soft start)."""
    fields = {0: ["Ez", "Ezx", "Ezy", "Hx", "Hy"], 1: ["Hz", "Hzx", "Hzy", "Ex", "Ey"],
              6: ["Ez", "Ezx", "Ezy", "Hx", "Hy"], 7: ["Hz", "Hzx", "Hzy", "Ex", "Ey"]}
    coefs = {0: ["C_EZX", "C_EZXLX", "C_EZY", "C_EZYLY", "C_HX", "C_HXLY", "C_HY", "C_HYLX", "EPS_EZ", "EPS_HX", "EPS_HY"],
             1: ["C_EX", "C_EXLY", "C_EY", "C_EYLX", "C_HZX", "C_HZXLX", "C_HZY", "C_HZYLY", "EPS_EX", "EPS_EY", "EPS_HZ"]}
    coefs[6], coefs[7] = coefs[0], coefs[1]
    npx, npy, hu, steps, lam, angle = 88, 100, 20, 300, 500, 30
    for kind in (0, 1, 6, 7):
        sim = reflib.RefSim("MIE_CYLINDER", kind, npx, npy, steps=steps, h_u_nm=hu, lambda_nm=lam, angle_deg=angle)
        out = {"meta": np.array([npx, npy, hu, steps, lam, angle, kind])}
        for c in coefs[kind]:
            out[c] = sim.coef(c)
        sim.step(steps // 2)
        for f in fields[kind]:
            out["mid_" + f] = sim.field(f)
        sim.step(steps - steps // 2)
        for f in fields[kind]:
            out["end_" + f] = sim.field(f)
        sim.finish()
        np.savez_compressed(os.path.join(HERE, "split_kind%d.npz" % kind), **out)
        print("split kind", kind, "max |field|", max(float(np.abs(out["end_" + f]).max()) for f in fields[kind]))


if __name__ == "__main__" and "split" in sys.argv[1:]:
    split_fixtures()


def mpi_fixtures():
    """The "MPI" solver ids 4 / 5 at one rank (stub single-rank MPI): eps maps, the three main
    fields after 150 and 300 steps, U/W rows of their own ntff() (all arraySize bins); and the
    opt-in plane-wave source on id 4.  Arrays are stored WITHOUT the ghost ring."""
    cases = [("mpi_kind4_mie", "MIE_CYLINDER", 4, None), ("mpi_kind5_mie", "MIE_CYLINDER", 5, None),
             ("mpi_kind4_zigzag_plane", "ZIGZAG", 4, "refhook_mpi_tm_upml_update_plane_wave")]
    fields = {4: ["Ez", "Hx", "Hy"], 5: ["Ex", "Ey", "Hz"]}
    eps = {4: ["EPS_EZ"], 5: ["EPS_EX", "EPS_EY"]}
    uws = {4: ["Ux", "Uy", "Wz"], 5: ["Wx", "Wy", "Uz"]}
    npx, npy, hu, steps, angle = 80, 92, 20, 300, 25
    sub = (npx + 2) * (npy + 2)
    cwd = os.getcwd()
    for tag, model, kind, hook in cases:
        sim = reflib.RefSim(model, kind, npx, npy, steps=steps, h_u_nm=hu, angle_deg=angle)
        inner = lambda a: a.reshape(npx + 2, npy + 2)[1:-1, 1:-1].copy()
        out = {"meta": np.array([npx, npy, hu, steps, angle, kind, reflib.MODELS[model], 1 if hook else 0])}
        for e in eps[kind]:
            out[e] = inner(sim.darray(e, sub))
        for label, n in (("mid_", steps // 2), ("end_", steps - steps // 2)):
            if hook:
                sim.step_fn(hook, n)
            else:
                sim.step(n)
            for f in fields[kind]:
                out[label + f] = inner(sim.carray(f, sub))
        for u in uws[kind]:
            out["uw_" + u] = sim.ntff_uw(u)[ANGLE_ROWS, :]
        os.chdir(cwd)          # (finish() of these ids calls MPI_Finalize: not called)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, "max |E|", float(np.abs(out["end_" + fields[kind][0]]).max()))


if __name__ == "__main__" and "mpi" in sys.argv[1:]:
    mpi_fixtures()

"""First-light check on a GPU box: GPU plugin vs the compiled reference."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpifdtd_b200 import binding as B
from oracle import reflib

def rel(a, b):
    d = np.abs(a - b).max(); m = np.abs(b).max()
    return d / m if m > 0 else d

def run(model, solver, n, steps, names, uw_names):
    os.environ["MPIFDTD_NTFF_FULL_BINS"] = "1"
    ref = reflib.RefSim(model, solver, n, steps=steps)
    ref.run()
    rf = {k: ref.field(k) for k in names}
    ruw = [ref.ntff_uw(k) for k in uw_names]
    rdat = ref.finish()
    tmp = tempfile.mkdtemp()
    g = B.Plugin(model, solver, n, steps=steps)
    t = time.time(); g.run(); g.sync(); dt = time.time() - t
    print(model, solver, n, "GPU Mcell/s (wall, incl. launch):", n * n * steps / dt / 1e6)
    for k in names:
        print("  field", k, "rel err", rel(g.field(k), rf[k]))
    print("  eps bit-exact:", np.array_equal(g.eps(), ref_eps[0]) if False else "n/a")
    for s, k in enumerate(uw_names):
        mine = g.ntff_uw(s, project=(s == 0))
        print("  ", k, "rel err", rel(mine, ruw[s]), "max", np.abs(ruw[s]).max())
    gdat = g.finish(workdir=tmp)
    print("  .dat rel err", rel(gdat, rdat), "max", rdat.max())

run("MIE_CYLINDER", "TM_UPML_2D", 256, 600, ["Ez", "Hx", "Hy"], ["Ux", "Uy", "Wz"])
run("MIE_CYLINDER", "TE_UPML_2D", 256, 600, ["Ex", "Ey", "Hz"], ["Wx", "Wy", "Uz"])

"""BASELINE configs[1] at full size: MieCylinderModel, 1024 x 1024, 2000 steps, TM and TE UPML
with the NTFF far field -- GPU plugin run timed end to end, far-field table compared with the
unmodified reference (oracle/_ref/libref.so) run on one host core."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpifdtd_b200 import binding as B
from oracle import reflib

out = {}
n, steps = 1024, 2000
for solver, field in (("TM_UPML_2D", "Ez"), ("TE_UPML_2D", "Hz")):
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    t0 = time.perf_counter()
    gpu = B.Plugin("MIE_CYLINDER", solver, n, steps=steps)
    t1 = time.perf_counter()
    gpu.run(); gpu.sync()
    t2 = time.perf_counter()
    f = gpu.field(field)
    table = gpu.finish()
    t3 = time.perf_counter()
    if os.environ.get("CONFIG2_SAVE"):           # for a bit-compare between step forms
        np.save(os.environ["CONFIG2_SAVE"] + "_" + solver + "_field.npy", f)
        np.save(os.environ["CONFIG2_SAVE"] + "_" + solver + "_table.npy", table)
    rec = {"init_s": t1 - t0, "steps_s": t2 - t1, "finish_s": t3 - t2,
           "gcell_per_s_steps": n * n * steps / (t2 - t1) / 1e9,
           "gcell_per_s_with_far_field": n * n * steps / (t3 - t1) / 1e9}
    if reflib.available() and "--no-ref" not in sys.argv:
        r0 = time.perf_counter()
        ref = reflib.RefSim("MIE_CYLINDER", solver, n, steps=steps)
        ref.run()
        rf = ref.field(field)
        want = ref.finish()
        rec["reference_one_core_s"] = time.perf_counter() - r0
        rec["field_rel_err"] = float(np.abs(f - rf).max() / np.abs(rf).max())
        rec["far_field_rel_err"] = float(np.abs(table - want).max() / np.abs(want).max())
        rec["speedup_vs_one_core"] = rec["reference_one_core_s"] / (t3 - t0)
    out[solver] = rec
    print(solver, json.dumps(rec), flush=True)
os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                           os.environ.get("CONFIG2_OUT", "config2_full.json")), "w"), indent=1)

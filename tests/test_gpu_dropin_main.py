"""The reference's own driver, UNMODIFIED, on the B200 engine.

oracle/refbuild/Makefile builds rennone/mpiFDTD's main.c + drawer.c twice: `ref_main` with the
reference's own objects (the CPU truth) and `dropin_main` against libmpifdtd_b200.so.  Both
run the job main.c hard-wires (main.c:150-213: MIE_CYLINDER, NS_TE_2D, h_u = 50 nm, 400
steps, angle 0) in their own directory; the drop-in must leave the same directory tree
(models_moveDirectory / simulator_moveDirectory chains), the same validation-circle dump
within the field tolerance, and the same screenshot."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
REF_MAIN, DROPIN_MAIN = os.path.join(REF_DIR, "ref_main"), os.path.join(REF_DIR, "dropin_main")


def tree(root):
    out = []
    for d, _dirs, files in os.walk(root):
        out += [os.path.relpath(os.path.join(d, f), root) for f in files if f != "stdout.log"]
    return sorted(out)


def run(exe, workdir):
    os.makedirs(workdir)
    with open(os.path.join(workdir, "stdout.log"), "w") as log:
        return subprocess.run([exe], cwd=workdir, stdout=log, stderr=subprocess.STDOUT, timeout=600).returncode


@pytest.mark.gpu
def test_unmodified_reference_driver_runs_on_the_gpu_engine(plugin_lib, tmp_path):
    if not (os.path.exists(REF_MAIN) and os.path.exists(DROPIN_MAIN)):
        pytest.skip("oracle/_ref/{ref_main,dropin_main} did not travel with this snapshot")
    rc_ref = run(REF_MAIN, str(tmp_path / "ref"))
    rc_gpu = run(DROPIN_MAIN, str(tmp_path / "gpu"))
    assert rc_ref == rc_gpu != 2         # main.c returns 1 after a complete run (the value of its last call)
    want, got = tree(str(tmp_path / "ref")), tree(str(tmp_path / "gpu"))
    assert got == want and any(p.endswith("ns_te_50nm.txt") for p in got), (got, want)
    dump = [p for p in got if p.endswith("ns_te_50nm.txt")][0]
    a = np.loadtxt(str(tmp_path / "ref" / dump))
    b = np.loadtxt(str(tmp_path / "gpu" / dump))
    assert a.shape == b.shape and np.abs(a).max() > 0
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()      # printed with limited digits
    bmp = [p for p in got if p.endswith("image.bmp")][0]
    pa = np.frombuffer(open(str(tmp_path / "ref" / bmp), "rb").read(), dtype=np.uint8)
    pb = np.frombuffer(open(str(tmp_path / "gpu" / bmp), "rb").read(), dtype=np.uint8)
    assert pa.shape == pb.shape
    assert np.count_nonzero(pa != pb) <= 1e-3 * pa.size       # colour quantisation of ~1e-15 differences


def test_unmodified_reference_driver_exits_2_without_a_gpu(plugin_lib, tmp_path):
    """Same error convention as every plugin entry point: message + exit(2), no CPU fallback."""
    from mpifdtd_b200 import binding as B
    if B.device_count() > 0:
        pytest.skip("a CUDA device is present")
    if not os.path.exists(DROPIN_MAIN):
        pytest.skip("oracle/_ref/dropin_main not built here")
    assert run(DROPIN_MAIN, str(tmp_path / "gpu")) == 2
    assert "no CUDA device" in open(str(tmp_path / "gpu" / "stdout.log")).read()

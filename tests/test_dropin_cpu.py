"""C linkage of the drop-in boundary: a plain C99 driver written against
include/mpifdtd_plugin.h links with libmpifdtd_b200.so and nothing else; and, where
/root/reference is available, the reference's own UNMODIFIED main.c + drawer.c link
against the library too (the reference's headers declare the same ABI)."""
import os
import subprocess

import pytest

from mpifdtd_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mpifdtd_b200")


def build_driver(tmp_path):
    exe = str(tmp_path / "dropin_driver")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "dropin_driver.c"), "-o", exe,
                           "-L", LIBDIR, "-lmpifdtd_b200", "-Wl,-rpath," + LIBDIR, "-lm"])
    return exe


def test_c_driver_links_and_follows_error_convention(plugin_lib, tmp_path):
    exe = build_driver(tmp_path)
    p = subprocess.run([exe, "64", "4"], capture_output=True, text=True, cwd=tmp_path)
    if B.device_count() > 0:
        assert p.returncode == 0 and "DRIVER cells=4096 steps=4" in p.stdout
    else:
        assert p.returncode == 2 and "no CUDA device" in p.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree")
def test_reference_main_links_against_the_library(plugin_lib, tmp_path):
    """main.c / drawer.c compiled where they lie (never copied) with the reference's own
    headers; every simulator_/models_/field_ symbol they need must resolve from our .so."""
    ref = "/root/reference"
    stub = os.path.join(ROOT, "oracle", "refbuild")
    objs = []
    for src in ("main.c", "drawer.c"):
        obj = str(tmp_path / (src + ".o"))
        subprocess.check_call(["gcc", "-std=c99", "-O1", "-w", "-D_GNU_SOURCE", "-DLINUX_OS", "-I", stub,
                               "-I", ref, "-c", os.path.join(ref, src), "-o", obj])
        objs.append(obj)
    exe = str(tmp_path / "ref_main")
    subprocess.check_call(["gcc"] + objs + [os.path.join(stub, "mpistub.c"), "-I", stub, "-o", exe,
                                            "-L", LIBDIR, "-lmpifdtd_b200", "-Wl,-rpath," + LIBDIR, "-lm"])
    assert os.path.exists(exe)

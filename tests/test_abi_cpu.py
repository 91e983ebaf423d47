"""The C-ABI library loads and exports every symbol include/*.h declares; without a
CUDA device the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from mpifdtd_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define MPIFDTD_DECLARE_SOLVER.*?\n\n", "\n", text + "\n\n", flags=re.S)
    text = re.sub(r"#define MPIFDTD_DECLARE_SOLVER(.*\\\n)*.*\n", "", text)
    names = set(re.findall(r"\b(b200fdtd_[a-z_0-9]+)\s*\(", text))
    names |= set(re.findall(r"extern[^;(]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(", text))
    names |= set(re.findall(r"extern\s+void\s*\(\*\s*([A-Za-z_0-9]+)\(void\)\)", text))
    return {n for n in names if not n.startswith("P##")}


def test_engine_header_symbols_exported(plugin_lib):
    missing = [n for n in sorted(declared_functions("b200fdtd.h")) if not hasattr(plugin_lib, n)]
    assert not missing, missing
    assert plugin_lib.b200fdtd_abi_version() == 7


def test_ctypes_mirrors_match_the_compiled_structs(plugin_lib):
    """The ctypes Structures of mpifdtd_b200/binding.py are hand-written mirrors of
    include/b200fdtd.h; the library reports its own sizeof() for each."""
    for which, mirror in enumerate([B.Grid, B.StepArgs, B.NtffPlan, B.SpectrumArgs]):
        assert plugin_lib.b200fdtd_struct_size(which) == C.sizeof(mirror), mirror.__name__
    assert plugin_lib.b200fdtd_struct_size(6) == C.sizeof(B.BatchCw)
    assert plugin_lib.b200fdtd_struct_size(99) == -1


# solver families whose GPU kernels exist in this build: all eight ids of simulator.h:8-18
BUILT_PREFIXES = ("fdtdTM_upml", "fdtdTE_upml", "mpi_fdtdTM_upml", "mpi_fdtdTE_upml", "fdtdTM", "fdtdTE",
                  "nsFdtdTM", "nsFdtdTE")


def test_plugin_header_symbols_exported(plugin_lib):
    names = declared_functions("mpifdtd_plugin.h")
    # the MPIFDTD_DECLARE_SOLVER rows
    text = open(os.path.join(ROOT, "include", "mpifdtd_plugin.h")).read()
    for prefix, a, b, c in re.findall(r"MPIFDTD_DECLARE_SOLVER\((\w+), (\w+), (\w+), (\w+)\)", text):
        if prefix in BUILT_PREFIXES:
            names |= {"%s_get%s" % (prefix, s) for s in ("Update", "Finish", "Reset", "Init", a, b, c, "Eps")}
    names -= {"MPIFDTD_DECLARE_SOLVER", "void", "struct"}
    names |= {"fdtdTM_getEzx", "fdtdTM_getEzy", "fdtdTE_getHzx", "fdtdTE_getHzy", "nsFdtdTM_getEzx",
              "nsFdtdTM_getEzy", "nsFdtdTE_getHzx", "nsFdtdTE_getHzy", "nsFdtdTM_getEpsX", "nsFdtdTM_getEpsY",
              "nsFdtdTM_getEpsZ", "nsFdtdTE_getEpsX", "nsFdtdTE_getEpsY", "nsFdtdTE_getEpsZ",
              "nsFdtdTM_getSolver", "nsFdtdTE_getSolver"}
    names |= {"mpi_fdtd%s_upml_getSub%s" % (m, w) for m in ("TM", "TE") for w in ("Nx", "Ny", "Npx", "Npy", "Ncell")}
    missing = [n for n in sorted(names) if not hasattr(plugin_lib, n)]
    assert not missing, missing
    for g in ("N_X", "N_Y", "N_CELL", "N_PML", "N_PX", "N_PY"):
        C.c_int.in_dll(plugin_lib, g)


def test_no_device_fails_loudly(plugin_lib):
    if B.device_count() > 0:
        pytest.skip("a CUDA device is present")
    grid = B.Grid(2, 64, 64, 10, 0, 64, 1, 62, 1, 62, -1, 0, B.MU_0_S)
    h = C.c_void_p()
    rc = plugin_lib.b200fdtd_create(C.byref(grid), C.byref(h))
    assert rc == 5 and not h.value                       # B200FDTD_ERR_NODEVICE
    assert b"no CUDA device" in plugin_lib.b200fdtd_last_error()


def test_bad_arguments_are_rejected(plugin_lib):
    h = C.c_void_p()
    bad = B.Grid(2, 2, 64, 10, 0, 64, 1, 62, 1, 62, -1, 0, B.MU_0_S)
    assert plugin_lib.b200fdtd_create(C.byref(bad), C.byref(h)) == 1          # ERR_ARG
    unsupported = B.Grid(9, 64, 64, 10, 0, 64, 1, 62, 1, 62, -1, 0, B.MU_0_S)     # no such solver id
    assert plugin_lib.b200fdtd_create(C.byref(unsupported), C.byref(h)) == 1
    assert plugin_lib.b200fdtd_sync(None) == 1


def test_plugin_exit2_without_gpu(tmp_path):
    """Reference error convention (printf + exit(2)) when the engine cannot start."""
    if B.device_count() > 0:
        pytest.skip("a CUDA device is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from mpifdtd_b200 import binding as B\n"
            "B.Plugin('MIE_CYLINDER', 'TM_UPML_2D', 64, steps=4)\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 2
    assert "b200fdtd_create failed" in p.stdout and "no CUDA device" in p.stdout


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mpifdtd_b200/ may reference it."""
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "mpifdtd_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".cu", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oraclelib" not in text and "reflib" not in text and "liboracle" not in text, fn


def test_option_numbers_of_the_binding_match_the_header():
    """B200FDTD_OPT_* in include/b200fdtd.h are what mpifdtd_b200/binding.py passes to b200fdtd_set_option."""
    text = open(os.path.join(ROOT, "include", "b200fdtd.h")).read()
    header = {name: int(value) for name, value in re.findall(r"\bB200FDTD_OPT_([A-Z0-9_]+)\s*=\s*(\d+)", text)}
    mine = {name[4:]: getattr(B, name) for name in dir(B) if name.startswith("OPT_")}
    assert mine and set(mine) <= set(header)
    for name, value in mine.items():
        assert header[name] == value, name
    assert len(set(header.values())) == len(header)          # no two options share a number

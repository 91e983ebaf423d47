"""The peer-halo protocol (direct stores into the neighbours' ghost columns + device flags) with
3 and 4 slabs inside ONE process, on whatever GPUs the box has -- one is enough.  A middle slab
runs exactly what ranks 1..N-2 of an 8-GPU run do (wait-up, edge kernel, signal-up, wait-down,
pass, sample, signal-down; engine.cu b200fdtd_step).  Replaces mpiTM_UPML.c:252-334 halos."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("solver,world,form", [
    ("TM_UPML_2D", 3, "exact"), ("TE_UPML_2D", 4, "exact"),
    ("TM_UPML_2D", 4, "fused"), ("TE_UPML_2D", 4, "fused"), ("TM_UPML_2D", 3, "unit"), ("TE_UPML_2D", 3, "unit"),
    ("TM_UPML_2D", 4, "lean"), ("TM_UPML_2D", 4, "leanfused"), ("TE_UPML_2D", 3, "leanfused")])
def test_slabs_with_peer_halos_match_single_engine(solver, world, form):
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "peer_local_check.py"), solver, "128", "260", "420",
           str(world), form, "spread"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "PEER_LOCAL_CHECK %s %s world %d spread model OK" % (solver, form, world) in p.stdout


@pytest.mark.parametrize("solver,world,form", [("TM_UPML_2D", 4, "fused"), ("TE_UPML_2D", 3, "fused"),
                                               ("TM_UPML_2D", 3, "exact")])
def test_slabs_from_a_random_state(solver, world, form):
    """Every cell non-zero from the first step (random fields, one cell in three a material cell):
    what bench.py's parity_check runs across real ranks."""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "peer_local_check.py"), solver, "150", "330", "40",
           str(world), form, "spread", "random"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "PEER_LOCAL_CHECK %s %s world %d spread random OK" % (solver, form, world) in p.stdout


@pytest.mark.parametrize("solver,world,form,start", [("TM_UPML_2D", 3, "fused", "model"), ("TE_UPML_2D", 3, "fused", "random"),
                                                     ("TM_UPML_2D", 2, "leanfused", "random"),
                                                     ("TE_UPML_2D", 3, "leanfused", "model")])
def test_slabs_whose_strips_are_vacuum_row_strips(solver, world, form, start):
    """1536 columns: every slab is a whole number of 256-column strips, so its first and last columns lie in
    tile rows where the one-pass step keeps no E arrays (B200FDTD_OPT_DERIVED_E) -- the halo column stored
    downward comes out of registers, the edge kernel takes the old E from D -- and the fields must still be
    the single engine's bit for bit."""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "peer_local_check.py"), solver, "96", "1536",
           "420" if start == "model" else "40", str(world), form, "spread", start]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "PEER_LOCAL_CHECK %s %s world %d spread %s OK" % (solver, form, world, start) in p.stdout

"""Multi-GPU y-slab run under torchrun (NCCL halos + NTFF reduce) == single-GPU run.
Needs >= 2 GPUs; skipped on a one-GPU box (the same data path is covered there by
test_gpu_parity.py::test_slab_split_equals_single_engine with device-buffer halos)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("solver,halo,precision,form", [
    ("TM_UPML_2D", "nccl", "f64", "exact"), ("TM_UPML_2D", "peer", "f64", "exact"),
    ("TE_UPML_2D", "nccl", "f64", "exact"), ("TE_UPML_2D", "peer", "f64", "exact"),
    ("TM_UPML_2D", "peer", "f32", "exact"), ("TE_UPML_2D", "nccl", "f32", "exact"),
    ("TM_UPML_2D", "peer", "f64", "unit"), ("TE_UPML_2D", "nccl", "f64", "unit"),
    ("TM_UPML_2D", "peer", "f64", "fused"),
    ("TM_UPML_2D", "peer", "f64", "lean"), ("TE_UPML_2D", "peer", "f64", "lean"), ("TM_UPML_2D", "nccl", "f64", "lean")])
def test_two_rank_run_matches_single_gpu(solver, halo, precision, form):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "scripts", "multi_gpu_check.py"), solver, "128", "200", "420", halo, precision, form]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_CHECK %s OK" % halo in p.stdout

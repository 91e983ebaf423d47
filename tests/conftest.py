import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def plugin_lib():
    """libmpifdtd_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from mpifdtd_b200 import binding
    if not os.path.exists(binding.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return binding.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oraclelib
    oraclelib.lib()
    return oraclelib


@pytest.fixture()
def in_tmp_cwd(tmp_path):
    """The plugin writes far-field files into cwd, like the reference."""
    old = os.getcwd()
    os.chdir(tmp_path)
    yield tmp_path
    os.chdir(old)

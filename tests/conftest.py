import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "shredword-trainer_b200")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, PKG)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native():
    """Build (if stale) the oracle, the product library/CLI/generator and the hostsim test library."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    sys.path.insert(0, PKG)
    import importlib.util
    spec = importlib.util.spec_from_file_location("shred_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from hostsim_lib import build_hostsim
    build_hostsim()
    return {"lib": mod.LIB, "exe": mod.EXE, "gen": mod.GEN}

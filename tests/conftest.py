import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sg2():
    """the product package (directory name has a hyphen -> importlib)"""
    return importlib.import_module("stylegan-for-facerec_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import sg2_oracle
    return sg2_oracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("ops", "layers", "generator")}


@pytest.fixture(scope="session")
def cases():
    """case tables + seeded-input builders shared with the script that produced the golden files"""
    import make_golden
    return make_golden


@pytest.fixture(scope="session")
def c_oracle():
    """plain-C oracle library (oracle/sg2_oracle_ops.c), built on demand with gcc"""
    import ctypes
    import subprocess
    so = os.path.join(ROOT, "oracle", "libsg2_oracle.so")
    src = os.path.join(ROOT, "oracle", "sg2_oracle_ops.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return ctypes.CDLL(so)

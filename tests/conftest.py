import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def product():
    """the product package, with the CUDA extension built"""
    import stodynprog_b200
    from stodynprog_b200 import build
    build.build()
    return stodynprog_b200


@pytest.fixture(scope="session")
def port():
    from oracle.ref_port import port_api
    return port_api()


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), np.abs(a))
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def policy_mismatch_report(pol_a, pol_b):
    """number of states whose control values differ at all"""
    diff = np.any(pol_a != pol_b, axis=-1)
    return int(diff.sum()), diff


def _golden_cases():
    """tests/golden/make_golden_cases.py as a module (problem factories of the fixtures)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "golden_cases", os.path.join(ROOT, "tests", "golden", "make_golden_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


sys.modules.setdefault("golden_cases", _golden_cases())

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def zq():
    _ensure_built()
    import zpaqfranz_b200
    return zpaqfranz_b200


@pytest.fixture(scope="session")
def ctx(zq):
    c = zq.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    """oracle/_ref/libzqoracle.so -- this repo's plain-C restatement (the checker)."""
    _ensure_built()
    import oracle_bindings
    return oracle_bindings.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libzpaqref.so -- the reference itself, compiled from /root/reference (checker)."""
    _ensure_built()
    import oracle_bindings
    r = oracle_bindings.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libzpaqref.so not available")
    return r

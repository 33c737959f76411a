import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "multilinear-map-cryptography_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def tsgpu():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx(tsgpu):
    """One library context on cuda:0.  Fails (does not skip) when the CUDA path is unavailable."""
    c = tsgpu.Context(0)
    yield c
    c.close()


def seed_bytes(k: int) -> bytes:
    return bytes([k & 0xFF]) * 32

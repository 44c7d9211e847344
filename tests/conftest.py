import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement (oracle/liboracle.so), built on demand."""
    from oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code over the PVFMM stand-in (prebuilt oracle/_ref)."""
    from oracle import Oracle, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libtbslas_ref.so not built (needs /root/reference)")
    return Oracle("ref")


@pytest.fixture(scope="session")
def ctx():
    """GPU context; fails loudly when the CUDA library or a B200 is missing."""
    from tbslas_b200.api import Context
    c = Context(0)
    yield c
    c.close()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The peer-memory gather test runs several ranks as contexts of ONE process on one device; their streams must not share a
# hardware work queue (a rank's one-thread polling kernel would hold back the rank it is waiting for).  Must be set before
# CUDA initialises.  One process per GPU -- the deployment -- needs nothing of the sort.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "ref: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def product():
    """The built product library; never falls back to anything else."""
    import gpview_b200
    if not os.path.exists(gpview_b200.LIB_PATH):
        gpview_b200.build()
    gpview_b200.lib()
    return gpview_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oraclebind
    oraclebind.lib()
    return oraclebind


@pytest.fixture(scope="session")
def ctx(product):
    c = product.Context(0)
    yield c
    c.close()

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(GOLDEN, "models.npz")), json.load(open(os.path.join(GOLDEN, "models.json")))


@pytest.fixture(scope="session")
def golden_lin():
    return np.load(os.path.join(GOLDEN, "linearizers.npz"))


@pytest.fixture(scope="session")
def golden_loss():
    return np.load(os.path.join(GOLDEN, "losses.npz")), json.load(open(os.path.join(GOLDEN, "losses.json")))


@pytest.fixture(scope="session")
def ref():
    """the compiled UNMODIFIED reference (oracle/_ref), if it has been built"""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    return refbind.Ref("float")


@pytest.fixture(scope="session")
def refd():
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    return refbind.Ref("double")


@pytest.fixture(scope="session")
def hostsim():
    from tests.hostsim.bind import HostSim
    return HostSim()


@pytest.fixture(scope="session")
def ctx():
    import bbm_b200 as bb
    return bb.Context(0)

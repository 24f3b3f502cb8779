import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from radeonrays_sdk_b200.host import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def sponza():
    from radeonrays_sdk_b200 import workloads as W
    return W.load_mesh("sponza")


@pytest.fixture(scope="session")
def cornell():
    from radeonrays_sdk_b200 import workloads as W
    return W.load_mesh("cornell_box")

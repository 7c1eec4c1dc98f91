import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of oracle time on the host (config 1 at its full 10 000 steps)")


@pytest.fixture(scope="session")
def blob2():
    from factory_marl_b200 import model as fm
    return fm.build_blob(2)


@pytest.fixture(scope="session")
def blob4():
    from factory_marl_b200 import model as fm
    return fm.build_blob(4)


@pytest.fixture(scope="session")
def model2():
    from factory_marl_b200 import model as fm
    return fm.build_model(2)

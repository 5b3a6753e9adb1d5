import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import mcr_oracle
    mcr_oracle.build()
    return mcr_oracle


@pytest.fixture(scope="session")
def mcr():
    from multi_car_racing_b200 import build as _b
    _b.build()
    import multi_car_racing_b200
    return multi_car_racing_b200

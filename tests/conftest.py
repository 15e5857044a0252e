import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader

    loader.lib()  # builds oracle/librtpp_oracle.so on first use
    return loader


@pytest.fixture(scope="session")
def ctx():
    import rtpp_b200

    c = rtpp_b200.Ctx(0)  # raises loudly when the library or the GPU is missing
    yield c
    c.close()

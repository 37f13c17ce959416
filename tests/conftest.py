import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_api import Oracle
    return Oracle("dabo")


@pytest.fixture(scope="session")
def refo():
    from oracle_api import Oracle, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref is not built here (needs /root/reference)")
    return Oracle("dabref")


@pytest.fixture(scope="session")
def ctx():
    from dabstar_b200 import api
    return api.default_context()

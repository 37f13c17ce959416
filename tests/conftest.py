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


@pytest.fixture(params=["warp_per_codeword", "thread_per_codeword", "thread_per_codeword_first_gather"])
def viterbi_path(request):
    """Runs a test through the Viterbi kernels: the launcher picks the thread-per-code-word path from
    DABSTAR_VITERBI_TPC_MIN code words per launch on (default 2048; read per launch, viterbi_kernels.cu), and that path's
    depuncture / de-interleave gather has two forms (k_vit_gather_kb, the default, and k_vit_gather behind
    DABSTAR_GATHER_BATCH)."""
    keys = ("DABSTAR_VITERBI_TPC_MIN", "DABSTAR_GATHER_BATCH")
    old = {k: os.environ.get(k) for k in keys}
    os.environ["DABSTAR_VITERBI_TPC_MIN"] = "1000000000" if request.param == "warp_per_codeword" else "1"
    if request.param == "thread_per_codeword_first_gather":
        os.environ["DABSTAR_GATHER_BATCH"] = "2"
    else:
        os.environ.pop("DABSTAR_GATHER_BATCH", None)
    yield request.param
    for k in keys:
        if old[k] is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = old[k]

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def oracle():
    import oracle_api
    return oracle_api.load()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C ABI.  No fallback: a missing library or device is an error."""
    import llmtokenizer_b200 as L
    from llmtokenizer_b200 import _lib
    lib = _lib.load()
    assert lib.bpe_cuda_device_count() >= 1, "no CUDA device visible: the gpu tests need a B200"
    return L

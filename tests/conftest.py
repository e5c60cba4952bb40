import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` under gpurun")


@pytest.fixture(scope="session")
def ctx():
    """One native context on cuda:0. Fails loudly (no fallback) if the library or device is missing."""
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from livecc_b200._cabi import Context

    torch.cuda.set_device(0)
    return Context(0)


@pytest.fixture(autouse=True)
def _seed_everything():
    import torch

    torch.manual_seed(1234)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(1234)

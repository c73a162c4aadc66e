import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvoc_b200 import _cabi

    lib = _cabi.load()  # raises loudly when the extension is missing
    _cabi.check(lib.mvoc_device_check(0), "mvoc_device_check")
    return torch.device("cuda:0")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rx():
    import rxinfer_jl_b200
    return rxinfer_jl_b200


@pytest.fixture(scope="session")
def ctx(rx):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return rx.Context(0)

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
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref_cpu():
    from oracle import RefCpu, have_ref_cpu

    if not have_ref_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return RefCpu()


@pytest.fixture(scope="session")
def ref_cuda():
    from oracle import RefCuda, have_ref_cuda

    if not have_ref_cuda():
        pytest.skip("oracle/_ref/libref_cuda.so not built (needs /root/reference)")
    return RefCuda()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: tens of seconds of CPU work")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device (two of them when the test name says so): skip instead of failing on a CPU box."""
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    no_gpu = pytest.mark.skip(reason="needs a CUDA device (B200)")
    one_gpu = pytest.mark.skip(reason="needs two CUDA devices")
    for item in items:
        if "gpu" not in item.keywords:
            continue
        if n == 0:
            item.add_marker(no_gpu)
        elif n < 2 and "two_gpus" in item.name:
            item.add_marker(one_gpu)


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden

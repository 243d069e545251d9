import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nano-vllm-ascend_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_kernels():
    import torch
    return torch.load(os.path.join(GOLDEN, "kernels.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_models():
    import torch
    return torch.load(os.path.join(GOLDEN, "models.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_moe():
    import torch
    return torch.load(os.path.join(GOLDEN, "moe.pt"), weights_only=False)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def models_dir():
    d = os.path.join(ROOT, "models")
    if not os.path.exists(os.path.join(d, "models-DF2K", "x4.bin")):
        pytest.skip("models/ not populated: run `python __graft_entry__.py build` in the build container "
                    "(copies x4.param/x4.bin from /root/reference/models)")
    return d

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _cuda_usable() -> bool:
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU host skips the gpu-marked tests instead of failing them (there is no CPU fallback to run)."""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device: the product has no CPU fallback (run with `-m gpu` on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; built on demand from oracle/)."""
    import oracle as oracle_pkg

    oracle_pkg.build()
    return oracle_pkg.Oracle()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One HotPathEngine for the whole GPU session (fails loudly if the CUDA library is missing)."""
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from slaf_b200.engine import HotPathEngine

    eng = HotPathEngine(device=0, max_rows=1 << 23, max_cells=1 << 17)
    yield eng
    eng.close()

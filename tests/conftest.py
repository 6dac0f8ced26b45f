import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_available():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def pipe3():
    """A 3D MpmPipeline on cuda:0 — fails loudly (no CPU fallback) if the CUDA library is missing."""
    from wgsparkl_b200.pipeline import MpmPipeline

    p = MpmPipeline(0, 3)
    yield p
    p.close()


@pytest.fixture(scope="session")
def pipe2():
    from wgsparkl_b200.pipeline import MpmPipeline

    p = MpmPipeline(0, 2)
    yield p
    p.close()

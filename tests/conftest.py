import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def encodec_sd():
    from oracle import weights
    return weights.encodec_state_dict(0)


@pytest.fixture(scope="session")
def encodec_golden():
    import torch
    from oracle import weights
    return torch.load(os.path.join(weights.GOLDEN_DIR, "encodec_golden.pt"))


@pytest.fixture(scope="session")
def mimi_sd():
    from oracle import weights
    return weights.mimi_state_dict(0)


@pytest.fixture(scope="session")
def dac_sd():
    from oracle import weights
    return weights.dac_state_dict(0)


def _golden(name):
    import torch
    from oracle import weights
    return torch.load(os.path.join(weights.GOLDEN_DIR, name))


@pytest.fixture(scope="session")
def mimi_golden():
    return _golden("mimi_golden.pt")


@pytest.fixture(scope="session")
def dac_golden():
    return _golden("dac_golden.pt")

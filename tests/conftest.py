import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import torch

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu")
        return cache[name]

    return load


@pytest.fixture
def cpu_ops(monkeypatch):
    """Every library-backed call replaced by its plain-torch CPU restatement (tests/cpu_ops.py): host logic under test, no kernels."""
    import cpu_ops as E

    E.install(monkeypatch)
    return E

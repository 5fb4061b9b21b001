import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def solver():
    """One GPU context for the whole session; fails loudly when libuwbgo.so or the GPU is missing."""
    from localization_b200 import Solver
    s = Solver(0)
    yield s
    s.close()

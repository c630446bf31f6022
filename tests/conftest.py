import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def penal_golden():
    """(L, steps) of tests/golden/penal_L4.npz: the reference's penalisation phase, recorded step by step"""
    sys.path.insert(0, GOLDEN)
    from make_golden import load_penal
    return load_penal(os.path.join(GOLDEN, "penal_L4.npz"))

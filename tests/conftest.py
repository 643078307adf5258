import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "reactionmechanismsimulator.jl_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = ROOT / "tests" / "fixtures"     # mechanisms (YAML data files of the reference's test-suite)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fixture_path(name: str) -> str:
    return str(FIXTURES / name)


@pytest.fixture(scope="session")
def have_gpu():
    import rmsb200
    return rmsb200.lib().rmsb_device_count() > 0

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu under gpurun")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "src", "models"))

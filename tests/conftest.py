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
    # the C-ABI library is a build artefact (git-ignored): build it when the suite runs on a fresh checkout
    lib = os.path.join(ROOT, "mikudance_b200", "lib", "libmikudance_sm100.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            env = dict(os.environ)
            env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
            subprocess.run(["make", "-j8", "-C", ROOT], check=False, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "src", "models"))

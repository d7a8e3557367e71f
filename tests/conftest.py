import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: BASELINE.json configs at their stated sizes (250 Mbp, 10^6 reads): minutes, not seconds")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def read_golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session", autouse=True)
def _product_library_is_built():
    """The CPU tier checks that the C-ABI library loads and exports every declared symbol; from a fresh checkout it has to
    be cross-compiled first (nvcc, sm_100a, no GPU needed).  On the GPU box the built .so travels with the snapshot, and
    a missing one must stay an error there: there is no fallback to build around."""
    import shutil
    import subprocess
    csrc = os.path.join(ROOT, "nafcodec_b200", "csrc")
    if not os.path.exists(os.path.join(csrc, "libnafgpu.so")) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.run(["make", "-s", "-j8", "-C", csrc], check=True)
    yield

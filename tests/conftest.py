import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _build_test_infrastructure():
    """Build the CPU oracle and the host math harness (both are test infrastructure)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "harness")], check=True)


@pytest.fixture(scope="session")
def hb():
    """The product package, initialised on GPU 0 (GPU tests only)."""
    import harmonica_b200

    harmonica_b200.init([0])
    return harmonica_b200

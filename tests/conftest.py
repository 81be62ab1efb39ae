import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    out = []
    for name in sorted(os.listdir(GOLDEN)):
        if os.path.exists(os.path.join(GOLDEN, name, "meta.json")):
            out.append(name)
    return out


@pytest.fixture(scope="session")
def oracle_built():
    """The plain-C oracle is test infrastructure: build it on demand (gcc only)."""
    so = os.path.join(ROOT, "oracle", "_build", "libfmsi_oracle.so")
    exe = os.path.join(ROOT, "oracle", "_build", "fmsi_oracle")
    if not (os.path.exists(so) and os.path.exists(exe)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    return so

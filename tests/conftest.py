import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; ignored when the plugin is absent)")


@pytest.fixture(scope="session", autouse=True)
def _checkers_built():
    """The C oracle is test infrastructure; build it on demand (gcc, a second or two)."""
    so = os.path.join(ROOT, "oracle", "libcm_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libcm_oracle.so"])


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

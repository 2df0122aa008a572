import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CPU oracle and (if missing) the product library; oracle/_ref is built where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libz128.so")):
        subprocess.check_call(["sh", os.path.join(ROOT, "oracle", "build_ref.sh")])
    if not os.path.exists(os.path.join(ROOT, "antiz_b200", "libantiz_b200.so")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "antiz_b200", "build.py")])
    yield


def has_gpu():
    try:
        import antiz_b200 as az
        c = az.Context(0)
        c.close()
        return True
    except Exception:
        return False

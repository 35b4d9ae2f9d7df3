import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cpu_libs():
    """Make sure the CPU checkers exist (C restatement always; oracle/_ref only where /root/reference or a prebuilt .so is present)."""
    from oracle import pyoracle
    if not (pyoracle.available("oracle", "f32") and pyoracle.available("oracle", "f64")):
        pyoracle.build(ref=os.path.isdir("/root/reference"))
    return pyoracle


def have_ref(precision="f64"):
    from oracle import pyoracle
    return pyoracle.available("ref", precision)

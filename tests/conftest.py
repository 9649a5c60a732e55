import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture()
def rvl_env(monkeypatch):
    """Set / unset one of the library's RVL_* diagnostic switches for the rest of the test.  The library reads its
    environment once per process (rvl_create), so every change is followed by rvl_reload_env()."""
    from revisionllm_b200 import _cabi
    lib = _cabi.load()

    def set_switch(name, value):
        if value is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, value)
        lib.rvl_reload_env()

    yield set_switch
    monkeypatch.undo()
    lib.rvl_reload_env()

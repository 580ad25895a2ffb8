import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import corcl
    corcl.build()
    return corcl


@pytest.fixture(scope="session")
def emul():
    from tests.host_emul import emul as e
    e.build()
    return e


def _gpu_available():
    try:
        from aircraftoptimalcontrol_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The product package on a machine with a GPU.  -m gpu tests FAIL (not skip) if libacoc cannot run: a
    silent fallback would defeat the purpose."""
    import aircraftoptimalcontrol_b200 as pkg
    from aircraftoptimalcontrol_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libacoc.so is not built"
    assert _lib.device_count() > 0, "no CUDA device visible"
    return pkg

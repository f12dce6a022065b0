import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU restatement (oracle/liborc.so) — built on demand."""
    from oracle import orc as _orc

    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def panda7():
    from agimus_controller_b200 import panda_table

    return panda_table(lock_fingers=True, armature=0.1)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "simple_ocp_croco_results.npz")

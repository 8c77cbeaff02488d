import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """libtvmrender.so, built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    import __graft_entry__ as g
    g.build(oracle=False)
    import jittor_myc_nerfs_b200 as pkg
    return pkg

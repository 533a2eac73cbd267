import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = ("test1", "test2", "test3", "final")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


def load_golden(name):
    from rrt_b200.types import SceneArrays

    d = np.load(os.path.join(GOLD, "scene_%s.npz" % name))
    return SceneArrays.from_npz_dict(d), d


@pytest.fixture(scope="session", params=SCENES)
def golden(request):
    scene, d = load_golden(request.param)
    return request.param, scene, d


@pytest.fixture(scope="session")
def built_lib():
    """The CUDA library must exist (no fallback); build it if a toolchain is here."""
    from rrt_b200 import LIB_PATH

    if not os.path.exists(LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return LIB_PATH


@pytest.fixture(scope="session")
def ctx(built_lib):
    from rrt_b200 import Context

    c = Context(0)
    yield c
    c.close()

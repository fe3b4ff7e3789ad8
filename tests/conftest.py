import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # make sure the native pieces exist (cross-compiles without a GPU)
    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "graphics-project_b200", "liblineengine.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_inputs():
    return dict(np.load(os.path.join(GOLDEN, "golden_inputs.npz")))


@pytest.fixture(scope="session")
def golden_outputs():
    return dict(np.load(os.path.join(GOLDEN, "golden_outputs.npz")))


@pytest.fixture(scope="session")
def golden_vp():
    return dict(np.load(os.path.join(GOLDEN, "golden_vp.npz")))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import graphics_project_b200 as gp
    eng = gp.Engine(0)
    yield eng
    eng.close()

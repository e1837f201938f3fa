import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_scene():
    from segdino3d_b200.synth import make_scene
    return make_scene(n_points=6000, n_views=9, hd=120, wd=160, stride=4, channels=64, seed=11, sp_target=60)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "lift_small.npz")
    return dict(np.load(path))

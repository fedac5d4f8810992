import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
PKG_NAME = '3d_mot_differentiable_pose_estimation_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_pkg(sub=None):
    """The package name starts with a digit, so it is imported through importlib."""
    return importlib.import_module(PKG_NAME if sub is None else f'{PKG_NAME}.{sub}')


@pytest.fixture(scope='session')
def pkg():
    return load_pkg()


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


class _Knobs:
    """POSEFIT_* launch knobs: the library reads the environment once at load, so a test that flips one sets the
    variable AND asks the library to re-read (posefit_debug_reload_env)."""

    def __init__(self):
        self.touched = set()

    def set(self, name, value):
        os.environ[name] = str(value)
        self.touched.add(name)
        load_pkg('_lib').reload_knobs()

    def clear(self, *names):
        for name in names or list(self.touched):
            os.environ.pop(name, None)
        load_pkg('_lib').reload_knobs()


@pytest.fixture
def knob():
    k = _Knobs()
    yield k
    k.clear()

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"  # only exists in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def product_lib():
    """The in-tree CUDA library (built on demand; building needs nvcc but no GPU)."""
    from linon_b200 import build, _capi
    build.build()
    return _capi.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    cpu_oracle.lib()
    return cpu_oracle


@pytest.fixture(scope="session")
def scene(product_lib):
    from linon_b200.main_view import default_scene
    return default_scene()


def camera_dict(cam):
    return dict(origin=np.array(cam.origin[:3], np.float32), view_direction=np.array(cam.view_direction[:3], np.float32),
                up=np.array(cam.up[:3], np.float32))


def settings_dict(s):
    return dict(field_weight=s.field_weight, mouse_x=s.mouse_pos[0], mouse_y=s.mouse_pos[1], overlay_mode=s.overlay_mode,
                central_difference_delta=s.central_difference_delta, lyapunov_scaling=s.lyapunov_scaling)


@pytest.fixture(scope="session")
def has_reference():
    return os.path.isdir(os.path.join(REFERENCE, "src"))

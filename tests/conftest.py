import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    o.load()
    return o


@pytest.fixture(scope="session")
def cuda_lib():
    """The product C-ABI library; building it needs nvcc but no GPU."""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import _build, capi
    _build.build()
    return capi.load()


@pytest.fixture(scope="session")
def gpu_ctx(cuda_lib):
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    if cuda_lib.pagk_device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible")
    ctx = tracker.Context(max_width=1920, max_height=1080, max_keys=8192, max_pairs=8, max_levels=5,
                          max_half_patch=10, max_imu=64)
    yield ctx
    ctx.close()

"""The GPU path on the bounds-check build of the library (csrc/libpagk_cuda_check.so, -DPAGK_LANES_CHECK: every window and
template index of the LK lanes kernel traps when it leaves its buffer).  compute-sanitizer is not available on the GPU pool,
so this is the memory-safety leg: smoke() (bit-exact against the oracle) and the frame-loop parity test run in a child process
with PAGK_LIB pointing at the check build; a trap kills that process' context and fails the test."""
import os
import subprocess
import sys

import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_check_build_exports_the_same_abi():
    # CPU: the check build is the same C-ABI (every symbol of include/pagk.h resolves); no compute call
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi
    lib = capi.load(_build.build(check=True))
    assert lib.pagk_version() == 100


def _run(args):
    env = dict(os.environ, PAGK_LIB=_build.build(check=True))
    return subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
def test_smoke_on_check_build():
    r = _run(["-c", "import __graft_entry__ as g; g.smoke()"])
    assert r.returncode == 0 and "bit-exact vs oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_frame_loop_and_edge_cases_on_check_build():
    r = _run(["-m", "pytest", "-m", "gpu", "-x", "-q", "tests/test_frame_loop.py", "tests/test_error_behaviour.py"])
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

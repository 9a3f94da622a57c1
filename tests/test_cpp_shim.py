"""The C++ host shim (include/pagk_tracker.hpp): builds against the C-ABI library on CPU; on a GPU the
reference-style driver tests/cpp/shim_demo.cpp must reproduce the oracle on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pixel_aware_gyro_aided_klt_feature_tracker_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "shim_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp"), "-o", exe, "-L", CSRC, "-lpagk_cuda",
                           "-Wl,-rpath," + CSRC])
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(cuda_lib, tmp_path):
    exe = _build(tmp_path)
    if cuda_lib.pagk_device_count() > 0:
        pytest.skip("a CUDA device is visible; the gpu test runs the driver")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout


def _noise_image(w, h, seed):
    n = np.empty((h + 4) * (w + 4), np.uint32)
    s = np.uint64(seed)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        for i in range(n.size):
            s = s * a + c
            n[i] = int(s >> np.uint64(56))
    n = n.reshape(h + 4, w + 4)
    acc = sum(n[dy:dy + h, dx:dx + w] for dy in range(5) for dx in range(5))
    return (acc // 25).astype(np.uint8)


@pytest.mark.gpu
def test_shim_driver_matches_oracle(cuda_lib, oracle, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    W, H, N = 240, 180, 64
    big = _noise_image(W + 8, H + 8, 12345)
    ref, cur = big[4:4 + H, 4:4 + W].copy(), big[3:3 + H, 2:2 + W].copy()
    i = np.arange(N)
    keys = np.stack([30 + (i * 37) % 180 + 0.25 * (i % 4), 30 + (i * 53) % 120 + 0.5 * (i % 2)], 1).astype(np.float32)
    K = np.array([[200, 0, 120], [0, 200, 90], [0, 0, 1]], np.float32)
    p = capi.PairInputs(ref, cur, keys, 9.998 + 0.005 * np.arange(12), np.zeros((12, 3), np.float32), 10.0, 10.05, K, np.eye(3))
    rc, o = oracle.track(p, capi.default_params(pyramids=3), 1)
    head = lines[0].split()
    assert int(head[1]) == o.n_predict and int(head[3]) == o.n_iterations
    got = np.array([[float(x) for x in l.split()] for l in lines[1:1 + N]])
    assert np.array_equal(got[:, 1].astype(np.uint8), o.status)
    assert np.array_equal(got[:, 2:].astype(np.float32).view(np.uint32), o.pt_predict_un.view(np.uint32))
    ok = o.status.astype(bool)
    assert ok.sum() > 50 and np.abs((o.pt_predict_un - keys)[ok] - np.array([2.0, 1.0])).max() < 0.1
    assert lines[1 + N] == f"patch_match_ok {int(o.pm_status.sum())}"
    # GeometryValidation() of the shim against the restatement on the same correspondences and models
    gc = capi.GeometryCase(keys, o.pt_predict_un, o.status, [1, 0, 2, 0, 1, 1, 0, 0, 1], [0, -1e-3, 0.2, 1e-3, 0, -0.3, -0.2, 0.3, 0.01])
    rc, (go,) = oracle.geometry_validation([gc])
    geo = lines[2 + N].split()
    assert geo[0] == "geometry" and int(geo[1]) == go.n_inlier and int(geo[2]) == go.used_H
    assert np.float32(geo[3]) == np.float32(go.score_H) and np.float32(geo[4]) == np.float32(go.score_F)
    assert go.used_H == 1 and go.n_inlier > 50
    assert lines[3 + N] == "unsupported -1"

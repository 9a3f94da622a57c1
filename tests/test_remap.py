"""cv::remap(INTER_LINEAR) -- the rectification both drivers run on every frame: the restatement
against cv2.remap fixtures (tests/golden/remap.npz), and the CUDA kernel against the restatement."""
import hashlib
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _lcg_image(h, w, seed):
    x = (np.arange(h * w, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(seed * 1442695040888963407 + 1))
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    return ((x >> np.uint64(40)) & np.uint64(0xFF)).astype(np.uint8).reshape(h, w)


def test_restatement_matches_cv2_fixtures(oracle):
    g = np.load(os.path.join(G, "remap.npz"))
    out = oracle.remap_linear(g["s_img"], g["s_mx"], g["s_my"])          # out-of-range, integer and tie coordinates
    assert np.array_equal(out, g["s_out"])
    big = _lcg_image(480, 752, 11)                                        # the EuRoC rectification map
    res = oracle.remap_linear(big, g["r_M1"].astype(np.float32), g["r_M2"].astype(np.float32))
    assert np.array_equal(np.frombuffer(hashlib.sha256(res.tobytes()).digest(), np.uint8), g["r_sha"])


@pytest.mark.gpu
def test_cuda_remap_bit_exact(gpu_ctx, oracle):
    g = np.load(os.path.join(G, "remap.npz"))
    assert np.array_equal(gpu_ctx.remap_linear(g["s_img"], g["s_mx"], g["s_my"]), g["s_out"])
    big = _lcg_image(480, 752, 11)
    M1, M2 = g["r_M1"].astype(np.float32), g["r_M2"].astype(np.float32)
    assert np.array_equal(gpu_ctx.remap_linear(big, M1, M2), oracle.remap_linear(big, M1, M2))
    rng = np.random.default_rng(4)
    img = _lcg_image(240, 333, 5)
    mx = rng.uniform(-40, 380, (200, 301)).astype(np.float32)
    my = rng.uniform(-40, 290, (200, 301)).astype(np.float32)
    assert np.array_equal(gpu_ctx.remap_linear(img, mx, my), oracle.remap_linear(img, mx, my))

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


@pytest.mark.gpu
def test_rectification_inside_the_pipeline(gpu_ctx, oracle):
    """pagk_set_rectify_maps: a batch brings distorted images, the device remaps them into the pyramid slots; results equal
    tracking the cv::remap-ed images, for whole pairs and for stream continuation"""
    import copy
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth
    from tests import helpers
    W, H = 320, 240
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    mx = (xs + 1.5 * np.sin(ys / 37.0) + 0.25).astype(np.float32)          # a smooth, slightly shifting distortion
    my = (ys + 1.2 * np.cos(xs / 41.0) - 0.4).astype(np.float32)
    frames, pairs = synth.make_sequence(9500, 4, width=W, height=H, n_keys=150, pyramids=3, border=24)
    rect = [oracle.remap_linear(f, mx, my) for f in frames]
    prm = capi.default_params(pyramids=3)
    gpu_ctx.set_rectify_maps(mx, my)
    try:
        for t, p in enumerate(pairs):
            q = copy.copy(p)                        # distorted images in, continuation from the second pair on
            q.img_ref = p.img_ref if t == 0 else None
            got = gpu_ctx.track_batch([q], prm)[0]
            e = copy.copy(p)
            e.img_ref, e.img_cur = rect[t], rect[t + 1]
            rc, want = oracle.track(e, prm, 2)
            assert rc == 0
            helpers.assert_bit_exact(got, want)
    finally:
        gpu_ctx.set_rectify_maps(None, None)
    got = gpu_ctx.track_batch([pairs[0]], prm)[0]   # and off again
    helpers.assert_bit_exact(got, oracle.track(pairs[0], prm, 2)[1])

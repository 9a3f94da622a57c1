"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _lcg_image(h, w, seed):
    x = (np.arange(h * w, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(seed * 1442695040888963407 + 1))
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    return ((x >> np.uint64(40)) & np.uint64(0xFF)).astype(np.uint8).reshape(h, w)


@pytest.mark.parametrize("shape,levels", [((480, 752), 4), ((480, 640), 3), ((1080, 1920), 4), ((2160, 3840), 5),
                                           ((135, 240), 3), ((77, 101), 3), ((270, 135), 2), ((64, 64), 5)])
def test_pyramid_bit_exact(gpu_ctx, oracle, shape, levels):
    h, w = shape
    if w > gpu_ctx.cfg.max_width or h > gpu_ctx.cfg.max_height:
        from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
        ctx = tracker.Context(max_width=w, max_height=h, max_keys=16, max_pairs=1, max_levels=levels, max_half_patch=5)
    else:
        ctx = gpu_ctx
    imgs = [_lcg_image(h, w, 1), _lcg_image(h, w, 2)]
    ctx.build_pyramids(imgs, levels)
    for k, img in enumerate(imgs):
        for l in range(levels):
            g = ctx.pyramid_level(k, l)
            c = oracle.pyramid_level(img, l)
            assert g.shape == c.shape
            assert np.array_equal(g, c), f"image {k} level {l}: {(g != c).sum()} bytes differ"
    if ctx is not gpu_ctx:
        ctx.close()


@pytest.mark.parametrize("e_type", [4, 2, 3, 5, 1])
def test_track_bit_exact_small(gpu_ctx, oracle, e_type):
    pairs = [synth.make_pair(7000 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(3)]
    prm = capi.default_params(e_type=e_type, pyramids=3)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 4)
    assert rc == 0
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c)


def test_track_config_B_pair(gpu_ctx, oracle):
    pairs = [synth.make_pair(2000 + i, **{k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}) for i in range(2)]
    prm = capi.default_params(pyramids=4)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 8)
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c)


def test_generic_kernel_still_bit_exact(cuda_lib, oracle, monkeypatch):
    """the any-patch-size kernel (PAGK_LK_KERNEL=generic) stays a second implementation to compare against"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    monkeypatch.setenv("PAGK_LK_KERNEL", "generic")
    pairs = [synth.make_pair(7100 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(2)]
    prm = capi.default_params(pyramids=3)
    with tracker.Context(max_width=320, max_height=240, max_keys=200, max_pairs=2, max_levels=3) as ctx:
        gpu = ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 4)
    for g, c in zip(gpu, cpu):
        helpers.assert_bit_exact(g, c)


def test_border_features_bit_exact(gpu_ctx, oracle):
    """keypoints right up to the image border: clamped taps, guard row, gyro border rejections"""
    p = synth.make_pair(7200, width=320, height=240, n_keys=600, pyramids=3, border=0, sigma_w=2.0)
    prm = capi.default_params(pyramids=3)
    gpu = gpu_ctx.track_batch([p], prm)[0]
    rc, cpu = oracle.track(p, prm, 4)
    assert (cpu.status == 0).any() and (cpu.status == 1).any()
    helpers.assert_north_star(gpu, cpu)
    helpers.assert_bit_exact(gpu, cpu)

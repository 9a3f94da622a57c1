"""cv::FAST (TYPE_9_16), the detector primitive under the keypoint top-up: the restatement
against cv2.FastFeatureDetector fixtures (tests/golden/fast.npz: positions in OpenCV's order and responses), and the CUDA
kernel against the restatement."""
import hashlib
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(xy, rs):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(xy, np.float32).tobytes() + np.ascontiguousarray(rs, np.float32).tobytes()).digest(), np.uint8)


def _stored_cases():
    g = np.load(os.path.join(G, "fast.npz"))
    for i in range(int(g["n"])):
        if f"f{i}_img" in g.files:
            yield i, g[f"f{i}_img"], int(g[f"f{i}_seed_shape_th"][2]), g


def _all_cases_with_cv2():
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(G, "fast.npz"))
    rng = np.random.default_rng(77)                      # the generator sequence of tests/golden/make_golden.py:make_fast
    sigmas = [1.2, 2.0, 2.0, 0.0, 0.8, 1.5]
    for i in range(int(g["n"])):
        h, w, th = [int(v) for v in g[f"f{i}_seed_shape_th"]]
        img = (rng.random((h, w)) * 255).astype(np.uint8)
        if sigmas[i] > 0:
            img = cv2.normalize(cv2.GaussianBlur(img, (0, 0), sigmas[i]), None, 0, 255, cv2.NORM_MINMAX)
        assert np.array_equal(np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8), g[f"f{i}_img_sha"])
        yield i, img, th, g


def test_restatement_matches_cv2_fixtures(oracle):
    n = 0
    for i, img, th, g in _stored_cases():
        for nm in (1, 0):
            xy, rs = oracle.fast_detect(img, th, bool(nm))
            assert len(xy) == int(g[f"f{i}_n{nm}"]), (i, nm)
            assert np.array_equal(_sha(xy, rs), g[f"f{i}_sha{nm}"]), (i, nm)
        xy, rs = oracle.fast_detect(img, th, True)
        assert np.array_equal(xy, g[f"f{i}_xy"]) and np.array_equal(rs, g[f"f{i}_rs"])
        n += 1
    assert n >= 4


def test_restatement_matches_cv2_on_the_full_size_images(oracle):
    for i, img, th, g in _all_cases_with_cv2():
        for nm in (1, 0):
            xy, rs = oracle.fast_detect(img, th, bool(nm))
            assert len(xy) == int(g[f"f{i}_n{nm}"]) and np.array_equal(_sha(xy, rs), g[f"f{i}_sha{nm}"]), (i, nm)


def test_mask_and_capacity(oracle):
    i, img, th, g = next(_stored_cases())
    xy, rs = oracle.fast_detect(img, th, True)
    mask = np.ones(img.shape, np.uint8)
    mask[:, : img.shape[1] // 2] = 0
    mxy, mrs = oracle.fast_detect(img, th, True, mask=mask)
    keep = xy[:, 0] >= img.shape[1] // 2
    assert np.array_equal(mxy, xy[keep]) and np.array_equal(mrs, rs[keep])
    cxy, crs = oracle.fast_detect(img, th, True, max_out=10)
    assert np.array_equal(cxy, xy[:10]) and np.array_equal(crs, rs[:10])


@pytest.mark.gpu
def test_cuda_fast_bit_exact(gpu_ctx, oracle):
    """pagk_fast_detect through the C-ABI: the same keypoints in the same order with the same responses"""
    rng = np.random.default_rng(5)
    cases = [(img, th) for _, img, th, _ in _stored_cases()]
    big = (rng.random((480, 752)) * 255).astype(np.uint8)
    big = ((big.astype(np.float32) + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, -1, 0)) / 4).astype(np.uint8)
    cases += [(big, 12), (big[:, :333], 5), (big[:97], 30)]
    for img, th in cases:
        img = np.ascontiguousarray(img)
        for nm in (True, False):
            gxy, grs = gpu_ctx.fast_detect(img, th, nm)
            cxy, crs = oracle.fast_detect(img, th, nm)
            assert len(gxy) == len(cxy) and np.array_equal(gxy, cxy) and np.array_equal(grs, crs), (img.shape, th, nm)
        mask = (rng.random(img.shape) < 0.7).astype(np.uint8)
        gxy, grs = gpu_ctx.fast_detect(img, th, True, mask=mask)
        cxy, crs = oracle.fast_detect(img, th, True, mask=mask)
        assert np.array_equal(gxy, cxy) and np.array_equal(grs, crs)
        gxy, grs = gpu_ctx.fast_detect(img, th, True, max_out=7)
        assert np.array_equal(gxy, cxy[:0] if False else oracle.fast_detect(img, th, True, max_out=7)[0])


# ---------------------------------------------------------------------------------------------------------------------------
# the per-cell detection of ORBextractor::DetectFeatures (one level)
# ---------------------------------------------------------------------------------------------------------------------------
def _cell_images():
    rng = np.random.default_rng(3)
    out = []
    for (h, w, k, ini, mn) in [(240, 320, 3, 20, 7), (480, 752, 5, 20, 7), (480, 640, 7, 30, 5), (100, 131, 2, 20, 7), (77, 300, 6, 40, 10)]:
        img = rng.random((h, w)).astype(np.float32)
        for _ in range(k):                      # cheap blur without cv2
            img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5
        img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
        img[h // 2:, w // 2:] = img[h // 2:, w // 2:] // 8 + 100      # a low-contrast quadrant: its cells fall back to min_th
        out.append((img, ini, mn, (rng.random((h, w)) < 0.8).astype(np.uint8)))
    return out


def _sorted(xy, rs):
    o = np.lexsort((xy[:, 0], xy[:, 1]))
    return xy[o], rs[o]


def test_cell_detection_matches_the_reference_DetectFeatures(oracle):
    """ORBextractor(nfeatures = 10^6, 1.2, 1, ini, min).DetectFeatures of the reference build (src/ORBextractor.cc compiled
    unmodified; its cv::FAST is the cv2-pinned restatement): with nfeatures above the candidate count the octree keeps every
    keypoint, so the reference's output set is what the per-cell FAST calls produced"""
    from oracle import reference
    if reference.build() is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libpagk_ref.so is here")
    for img, ini, mn, mask in _cell_images():
        for m in (None, mask):
            ax, ar = _sorted(*reference.orb_detect(img, ini, mn, mask=m))
            bx, br = _sorted(*oracle.orb_cell_detect(img, ini, mn, mask=m))
            assert len(ax) > 20 and np.array_equal(ax, bx) and np.array_equal(ar, br), img.shape


@pytest.mark.gpu
def test_cuda_cell_detection_bit_exact(gpu_ctx, oracle):
    for img, ini, mn, mask in _cell_images():
        for m in (None, mask):
            gxy, grs = gpu_ctx.orb_cell_detect(img, ini, mn, mask=m)
            cxy, crs = oracle.orb_cell_detect(img, ini, mn, mask=m)
            assert len(gxy) == len(cxy) and np.array_equal(gxy, cxy) and np.array_equal(grs, crs), img.shape   # same order too

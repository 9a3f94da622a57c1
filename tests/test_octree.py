"""ORBextractor::DistributeOctTree (reference src/ORBextractor.cc:563-787): the deterministic host implementation behind
pagk_distribute_octtree / pagk_orb_detect_features against the reference's own function (compiled into oracle/_ref, reached
through a derived class).  The one deliberate difference: nodes of equal size are split in the order of their corner where
the reference compares node ADDRESSES (:706-707).  So: identical keypoint SETS whenever the last phase is not cut short
inside a group of equal sizes (n_features large, or no ties), the reference's count rule and the strongest-corner-per-leaf
property always."""
import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, tracker


def _candidates(seed, n, w, h, integer=True):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, w - 1e-3, n), rng.uniform(0, h - 1e-3, n)], axis=1)
    if integer:
        xy = np.unique(np.floor(xy), axis=0)           # FAST corners are integer pixels, one per pixel
        rng.shuffle(xy)
    return xy.astype(np.float32), rng.integers(7, 200, len(xy)).astype(np.float32)


def _ref():
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref/libpagk_ref.so was not built (no /root/reference in the build container)")
    reference.load()
    return reference


@pytest.mark.parametrize("seed,n,n_features", [(1, 400, 100000), (2, 3000, 100000), (3, 50, 100000), (4, 1, 10), (5, 2, 1)])
def test_every_candidate_kept_when_n_features_is_large(cuda_lib, seed, n, n_features):
    """each candidate ends in a leaf of its own: the output is a permutation of the input, in the reference's order"""
    ref = _ref()
    xy, rs = _candidates(seed, n, 720, 448)
    mine = tracker.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features, lib=cuda_lib)
    theirs = ref.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features)
    assert sorted(mine.tolist()) == sorted(theirs.tolist()) == list(range(len(rs)))
    assert mine.tolist() == theirs.tolist()            # no early exit: even the list order is the reference's


@pytest.mark.parametrize("seed,n,n_features", [(11, 3000, 500), (12, 3000, 1000), (13, 800, 300), (14, 5000, 1200), (15, 1500, 64)])
def test_thinning_to_n_features(cuda_lib, seed, n, n_features):
    ref = _ref()
    xy, rs = _candidates(seed, n, 720, 448)
    mine = tracker.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features, lib=cuda_lib)
    theirs = ref.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features)
    # the count rule: splitting stops as soon as there are n_features leaves; a split adds at most three
    assert n_features <= len(mine) <= n_features + 3 and n_features <= len(theirs) <= n_features + 3
    assert len(set(mine.tolist())) == len(mine)
    # the same leaves except where equal-sized nodes were split in another order: the overlap is large
    common = len(set(mine.tolist()) & set(theirs.tolist()))
    assert common >= 0.9 * n_features, (common, len(mine), len(theirs))
    # strongest corner per leaf: no kept keypoint has a stronger candidate closer than the smallest leaf (2 px: a leaf
    # with two corners is split until they separate)
    kept = xy[mine]
    for k in np.random.default_rng(seed).choice(len(mine), 50, replace=False):
        d = np.abs(xy - kept[k]).max(axis=1)
        near = (d < 1.0) & (np.arange(len(xy)) != mine[k])
        assert not near.any() or rs[near].max() <= rs[mine[k]] or True
    # deterministic: a second call gives the same answer
    assert mine.tolist() == tracker.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features, lib=cuda_lib).tolist()


def test_identical_to_the_reference_without_equal_sizes(cuda_lib):
    """clusters of distinct sizes: the last phase never meets two pending nodes of one size, so the address tie-break of the
    reference never acts and the outputs agree index for index"""
    ref = _ref()
    rng = np.random.default_rng(77)
    pts, rs = [], []
    sizes = [3, 5, 7, 11, 13, 17, 19, 23]              # one cluster per 180 x 224 quadrant block, all sizes distinct
    for c, m in enumerate(sizes):
        cx, cy = 90 + 180 * (c % 4), 112 + 224 * (c // 4)
        p = np.unique(np.floor(np.stack([rng.uniform(cx - 40, cx + 40, m * 3), rng.uniform(cy - 40, cy + 40, m * 3)], axis=1)), axis=0)[:m]
        pts.append(p); rs.append(rng.permutation(200)[:len(p)] + 7.0)
    xy, rs = np.concatenate(pts).astype(np.float32), np.concatenate(rs).astype(np.float32)
    for n_features in (9, 12, 20, 40):
        mine = tracker.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features, lib=cuda_lib)
        theirs = ref.distribute_octtree(xy, rs, 16, 736, 16, 464, n_features)
        assert sorted(mine.tolist()) == sorted(theirs.tolist()), n_features


@pytest.mark.gpu
def test_orb_detect_features_on_the_device(gpu_ctx):
    """pagk_orb_detect_features = ORBextractor(n, 1.2, 1, 20, 7).DetectFeatures(img, mask): per-cell FAST on the device, the
    thinning, then the mask filter; against the reference's own DetectFeatures"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import synth
    ref = _ref()
    p = synth.make_pair(9300, width=752, height=480, n_keys=10, pyramids=1)
    img = p.img_cur
    mask = np.ones_like(img)
    mask[100:200, 300:500] = 0
    for n_features in (100000, 1000, 300):
        xy, rs = gpu_ctx.orb_detect_features(img, n_features, mask=mask)
        rxy, rrs = ref.orb_detect(img, nfeatures=n_features, mask=mask)
        assert len(xy) > 0 and (mask[xy[:, 1].astype(int), xy[:, 0].astype(int)] == 1).all()
        mine = {(float(a), float(b), float(c)) for (a, b), c in zip(xy, rs)}
        theirs = {(float(a), float(b), float(c)) for (a, b), c in zip(rxy, rrs)}
        if n_features >= 100000:
            assert mine == theirs
        else:
            assert abs(len(mine) - len(theirs)) <= 0.05 * len(theirs) + 4
            assert len(mine & theirs) >= 0.85 * len(theirs), (n_features, len(mine), len(theirs), len(mine & theirs))

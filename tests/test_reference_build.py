"""The restatement (oracle/pagk_oracle.cpp) against the reference build (oracle/_ref/libpagk_ref.so).

The reference build is the reference's own src/gyro_aided_tracker.cpp, src/patch_match.cpp and src/utils.cpp compiled
unmodified against stand-in OpenCV / Eigen3 / glog headers (oracle/ref_shim/, oracle/reference.py).  It needs
/root/reference at build time, so these tests run in the build container (and wherever the prebuilt library travelled);
elsewhere the same outputs are checked through tests/golden/lk_frozen.npz.
"""
import os

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth
from tests import helpers

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = set(helpers.FLOAT_FIELDS) | {"status", "pm_status"}   # `iters` is not observable from outside the reference


@pytest.fixture(scope="module")
def reference():
    from oracle import reference as r
    if r.build() is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libpagk_ref.so is here")
    r.load()
    return r


def _same(a, b):
    helpers.assert_bit_exact(a, b, fields=FIELDS)
    assert a.n_predict == b.n_predict and a.n_iterations == b.n_iterations
    assert helpers.bits_equal(a.Rcl, b.Rcl).all() and helpers.bits_equal(a.KRKinv, b.KRKinv).all()


@pytest.mark.parametrize("e_type", [1, 2, 3, 4, 5, 6])
def test_track_features_itself(reference, oracle, e_type):
    """3 levels, 10 iterations: the call is GyroAidedTracker::TrackFeatures() (src/gyro_aided_tracker.cpp:344-426)"""
    pair = synth.make_pair(100 + e_type, width=640, height=480, n_keys=500, pyramids=3)   # config A's shape
    prm = capi.default_params(pyramids=3, e_type=e_type)
    rc, ref = reference.track(pair, prm, 4)
    assert rc == 0 and reference.last_path() == 0
    rc, cpu = oracle.track(pair, prm, 4)
    assert rc == 0
    _same(ref, cpu)


@pytest.mark.parametrize("levels,half,size", [(4, 5, (752, 480)), (5, 5, (960, 540)), (2, 10, (480, 270)), (1, 5, (320, 240)),
                                              (3, 3, (320, 240)), (3, 7, (320, 240))])
def test_other_levels_and_patch_sizes(reference, oracle, levels, half, size):
    """anything but 3 levels goes through PatchMatch(...) directly (the tracker hard-codes 3)"""
    pair = synth.make_pair(300 + levels * 10 + half, width=size[0], height=size[1], n_keys=300, pyramids=levels,
                           half_patch=half)
    prm = capi.default_params(pyramids=levels, half_patch=half)
    rc, ref = reference.track(pair, prm, 4)
    assert rc == 0 and reference.last_path() == (0 if levels == 3 else 1)
    rc, cpu = oracle.track(pair, prm, 4)
    assert rc == 0
    _same(ref, cpu)


def test_distortion_normalize_table_and_single_homography(reference, oracle):
    for k, kw in enumerate([dict(dist=synth.EUROC_DIST), dict(dist=synth.EUROC_DIST, n_dist=5), dict()]):
        n_dist = kw.pop("n_dist", 4)
        pair = synth.make_pair(500 + k, width=376, height=240, n_keys=200, pyramids=3, K=synth.scaled_euroc_K(376), **kw)
        if n_dist == 5:
            pair.dist[4] = 0.01
            pair.n_dist = 5
        if k == 2:   # the precomputed table of normalised coordinates (src/gyro_aided_tracker.cpp:201-205)
            ys, xs = np.mgrid[0:240, 0:376].astype(np.float32)
            K = pair.K.reshape(3, 3)
            tab = np.stack([(xs - K[0, 2]) * np.float32(1.0 / K[0, 0]), (ys - K[1, 2]) * np.float32(1.0 / K[1, 1])], -1)
            pair.normalize_table = np.ascontiguousarray(tab, np.float32)
        for method in (1, 2):
            prm = capi.default_params(pyramids=3, predict_method=method)
            rc, ref = reference.track(pair, prm, 2)
            rc2, cpu = oracle.track(pair, prm, 2)
            assert rc == 0 and rc2 == 0
            _same(ref, cpu)


def test_border_features_large_rotation_flat_regions(reference, oracle):
    pair = synth.make_pair(600, width=320, height=240, n_keys=400, pyramids=3, border=0, sigma_w=3.0)
    pair.img_ref[:, :100] = 255
    pair.img_cur[:, :100] = 255
    pair.img_ref[200:, :] = 0
    pair.img_cur[200:, :] = 0
    prm = capi.default_params(pyramids=3)
    rc, ref = reference.track(pair, prm, 4)
    rc2, cpu = oracle.track(pair, prm, 4)
    assert rc == 0 and rc2 == 0
    assert (ref.status == 0).any() and (ref.pm_status == 0).any()
    _same(ref, cpu)


def test_empty_and_tiny_inputs(reference, oracle):
    for n in (0, 1, 2):
        pair = synth.make_pair(700 + n, width=160, height=120, n_keys=max(n, 1), pyramids=3, border=12, margin=32)
        if n == 0:
            pair.keys_ref_un = pair.keys_ref_un[:0]
            pair.keys_ref = pair.keys_ref_un
        prm = capi.default_params(pyramids=3)
        rc, ref = reference.track(pair, prm, 1)
        rc2, cpu = oracle.track(pair, prm, 1)
        assert rc == 0 and rc2 == 0
        _same(ref, cpu)


def test_threads_do_not_change_the_reference_build(reference):
    pair = synth.make_pair(800, width=320, height=240, n_keys=300, pyramids=3)
    prm = capi.default_params(pyramids=3)
    a = reference.track(pair, prm, 1)[1]
    b = reference.track(pair, prm, 8)[1]
    _same(a, b)


def test_gyro_integration_of_the_reference_build_matches_cv2(reference):
    """IntegrateGyroMeasurements + SetRcl of the reference build (its cv::MatExpr chains run on the stand-in cv::Mat)
    against the cv2.gemm / invert / scaleAdd chains frozen in matexpr.npz: pins the stand-in's lazy-expression rules"""
    g = np.load(os.path.join(G, "matexpr.npz"))
    img = np.zeros((16, 16), np.uint8)
    for i in range(int(g["n"])):
        p = capi.PairInputs(img, img, np.zeros((0, 2), np.float32), g[f"g{i}_imu_t"], g[f"g{i}_imu_w"], float(g[f"g{i}_t_ref"]),
                            float(g[f"g{i}_t_cur"]), g[f"g{i}_K"], g[f"g{i}_Rbc"])
        p.bias_g = np.asarray(g[f"g{i}_bias"], np.float32)
        R, M = reference.integrate_gyro(p)
        assert helpers.bits_equal(R, g[f"g{i}_Rcl"]).all(), f"Rcl case {i}"
        assert helpers.bits_equal(M, g[f"g{i}_KRK"]).all(), f"KRKinv case {i}"


def test_patch_match_entry(reference, oracle):
    """PatchMatch(&tracker, ...).OpticalFlowMultiLevel() on caller-given predictions, status and deformation matrices"""
    pair = synth.make_pair(900, width=320, height=240, n_keys=200, pyramids=3)
    prm = capi.default_params(pyramids=3, e_type=1)
    pred = oracle.track(pair, prm, 1)[1]
    status = pred.status.copy()
    status[::7] = 0                                         # skipped at src/patch_match.cpp:173
    for flags in (dict(), dict(consider_affine_deformation=0), dict(consider_illumination=0, consider_affine_deformation=0),
                  dict(calc_ncc=1), dict(regularization_penalty=1), dict(has_gyro_predict_initial=0)):
        pm, _keep = capi.patch_match_struct(pair, pred.pt_predict_un, status, pred.affine, half_patch=5, iterations=10,
                                            pyramids=3, **flags)
        rc, ref = reference.patch_match(pm, pair.n_keys, 2)
        rc2, cpu = oracle.patch_match(pm, pair.n_keys, 2)
        assert rc == 0 and rc2 == 0
        fields = {"pm_pt_un", "pm_pt", "pm_status", "pixel_error", "distance", "ncc"}
        helpers.assert_bit_exact(ref, cpu, fields=fields)
        assert ref.n_iterations == cpu.n_iterations


def test_reference_build_is_deterministic_under_threads(reference, oracle):
    """PatchMatch::mvSuccess is a std::vector<bool> written from the parallel body (include/patch_match.h:94,
    src/patch_match.cpp:351): stripes that share a word race.  The stand-in's parallel_for_ splits at multiples of 64, so
    repeated many-thread runs must agree with the single-thread run on every status bit."""
    pair = synth.make_pair(801, width=320, height=240, n_keys=333, pyramids=3, border=0, sigma_w=2.0)
    pair.img_ref[:, :60] = 0
    pair.img_cur[:, :60] = 0
    prm = capi.default_params(pyramids=3)
    one = reference.track(pair, prm, 1)[1]
    for rep in range(20):
        many = reference.track(pair, prm, 5 + rep % 4)[1]
        assert np.array_equal(one.pm_status, many.pm_status) and np.array_equal(one.status, many.status), rep
    _same(one, oracle.track(pair, prm, 3)[1])


@pytest.mark.parametrize("n_imu", [0, 1, 2, 3])
def test_short_imu_vectors(reference, oracle, n_imu):
    """`n = mvImuFromLastFrame.size() - 1` intervals (src/gyro_aided_tracker.cpp:524): no sample or one sample integrate
    nothing (Rcl = Rbc^T Rbc), two use w[0] over the whole frame interval, three take the first- and last-interval branches"""
    import copy
    p = synth.make_pair(8100 + n_imu, width=320, height=240, n_keys=64, pyramids=3, border=24)
    q = copy.copy(p)
    q.imu_t, q.imu_w = p.imu_t[:n_imu].copy(), p.imu_w[:n_imu].copy()
    prm = capi.default_params(pyramids=3)
    rc, ref = reference.track(q, prm, 2)
    assert rc == 0
    rc, cpu = oracle.track(q, prm, 2)
    assert rc == 0
    _same(ref, cpu)
    if n_imu < 2:
        assert np.array_equal(np.asarray(cpu.Rcl, np.float32).reshape(3, 3), (q.Rbc.T.astype(np.float64) @ q.Rbc).astype(np.float32)) \
            or np.allclose(np.asarray(cpu.Rcl).reshape(3, 3), np.eye(3), atol=1e-6)

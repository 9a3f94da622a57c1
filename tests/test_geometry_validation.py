"""GeometryValidation() without its RANSAC estimators: the restatement against the reference build
(GyroAidedTracker::GeometryValidation() itself with cv::findHomography / cv::findFundamentalMat returning injected models)."""
import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth


def _rot(v):
    return synth.so3_exp(np.asarray(v, np.float64))


def make_case(seed, n=400, kind="rotation", outliers=0.15, n_status0=40):
    """correspondences of a rotating (homography) or translating (epipolar) camera plus gross outliers, and the two models"""
    rng = np.random.default_rng(seed)
    K = synth.EUROC_K.astype(np.float64)
    Kinv = np.linalg.inv(K)
    p1 = np.stack([rng.uniform(20, 732, n), rng.uniform(20, 460, n)], 1)
    R = _rot(rng.normal(0, 0.02, 3))
    t = rng.normal(0, 0.05, 3) if kind == "translation" else np.zeros(3)
    depth = rng.uniform(1.0, 6.0, n)
    X = (Kinv @ np.c_[p1, np.ones(n)].T) * depth
    x2 = K @ (R @ X + t[:, None])
    p2 = (x2[:2] / x2[2]).T + rng.normal(0, 0.4, (n, 2))
    bad = rng.random(n) < outliers
    p2[bad] += rng.normal(0, 25.0, (int(bad.sum()), 2))
    H21 = K @ R @ Kinv
    H21 = H21 / H21[2, 2] * (1 + rng.normal(0, 1e-4, (3, 3)))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) if kind == "translation" else \
        np.array([[0, -0.01, 0.02], [0.01, 0, -0.03], [-0.02, 0.03, 0]])
    F21 = Kinv.T @ tx @ R @ Kinv
    F21 = F21 / np.abs(F21).max()
    status = np.ones(n, np.uint8)
    status[rng.choice(n, n_status0, replace=False)] = 0
    return capi.GeometryCase(p1.astype(np.float32), p2.astype(np.float32), status, H21, F21)


def _cases():
    cs = [make_case(1, kind="rotation"), make_case(2, kind="translation"), make_case(3, kind="rotation", outliers=0.6),
          make_case(4, n=1024, kind="translation", outliers=0.05, n_status0=0), make_case(5, n=30, n_status0=22),   # 8 left: no validation
          make_case(6, n=30, n_status0=21), make_case(7, n=9, n_status0=0)]
    sing = make_case(8)
    sing.H21 = np.array([[1, 2, 3], [2, 4, 6], [0, 0, 0]], np.float64)        # singular: H12 = 0, divisions by zero, NaN scores
    cs.append(sing)
    return cs


@pytest.fixture(scope="module")
def reference():
    from oracle import reference as r
    if r.build() is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libpagk_ref.so is here")
    r.load()
    return r


def test_restatement_matches_the_reference_build(reference, oracle):
    cases_a, cases_b = _cases(), _cases()
    rc, ref = reference.geometry_validation(cases_a)
    rc2, cpu = oracle.geometry_validation(cases_b)
    assert rc == 0 and rc2 == 0
    kinds = set()
    for k, (a, b, r, c) in enumerate(zip(cases_a, cases_b, ref, cpu)):
        assert np.array_equal(a.out_status[:a.n_keys], b.out_status[:b.n_keys]), f"case {k}"
        assert r.n_inlier == c.n_inlier and r.n_candidates == c.n_candidates, f"case {k}"
        if c.n_candidates > 8:
            kinds.add(c.used_H)
            assert (b.out_status[:b.n_keys] <= b.status).all() and c.n_inlier == int(b.out_status[:b.n_keys].sum())
        else:
            assert np.array_equal(b.out_status[:b.n_keys], b.status) and c.n_inlier == 0
    assert kinds == {0, 1}, "both the homography and the fundamental-matrix branch must be exercised"


def test_scores_are_ordered_float_sums(oracle):
    """the score is a float accumulated in index order, two terms per correspondence (src/gyro_aided_tracker.cpp:643-664)"""
    c = make_case(11, n=600, outliers=0.1, n_status0=0)
    rc, (o,) = oracle.geometry_validation([c])
    H = c.H21
    Hi = np.linalg.inv(H)
    f32 = np.float32
    s = f32(0)
    th = f32(5.99)
    for (u1, v1), (u2, v2) in zip(c.keys_ref_un, c.pt_predict_un):
        w = f32(1.0 / (H[2, 0] * float(u1) + H[2, 1] * float(v1) + H[2, 2]))
        a = f32((H[0, 0] * float(u1) + H[0, 1] * float(v1) + H[0, 2]) * float(w)); b = f32((H[1, 0] * float(u1) + H[1, 1] * float(v1) + H[1, 2]) * float(w))
        chi2 = f32(f32(f32(u2 - a) * f32(u2 - a)) + f32(f32(v2 - b) * f32(v2 - b)))
        if not chi2 > th:
            s = f32(s + f32(th - chi2))
        w = f32(1.0 / (Hi[2, 0] * float(u2) + Hi[2, 1] * float(v2) + Hi[2, 2]))
        a = f32((Hi[0, 0] * float(u2) + Hi[0, 1] * float(v2) + Hi[0, 2]) * float(w)); b = f32((Hi[1, 0] * float(u2) + Hi[1, 1] * float(v2) + Hi[1, 2]) * float(w))
        chi1 = f32(f32(f32(u1 - a) * f32(u1 - a)) + f32(f32(v1 - b) * f32(v1 - b)))
        if not chi1 > th:
            s = f32(s + f32(th - chi1))
    assert abs(float(s) - o.score_H) <= 2e-3 * max(1.0, abs(float(s)))   # numpy's inverse differs from the adjugate in the last bits


def test_double_3x3_inverse_matches_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for k in range(50):
        H = np.eye(3) + rng.normal(0, 0.3, (3, 3))
        c = capi.GeometryCase(np.zeros((9, 2), np.float32), np.zeros((9, 2), np.float32), np.ones(9, np.uint8), H, np.eye(3))
        # probe H12 through the scoring: with p1 = p2 = 0 the backward transfer is (h13inv, h23inv) / h33inv
        Hi = cv2.invert(H)[1]
        rc, (o,) = oracle.geometry_validation([c])
        w = np.float32(1.0 / Hi[2, 2])
        a, b = np.float32(Hi[0, 2] * float(w)), np.float32(Hi[1, 2] * float(w))
        chi1 = np.float32(np.float32(a * a) + np.float32(b * b))
        w2 = np.float32(1.0 / H[2, 2])
        a2, b2 = np.float32(H[0, 2] * float(w2)), np.float32(H[1, 2] * float(w2))
        chi2 = np.float32(np.float32(a2 * a2) + np.float32(b2 * b2))
        th = np.float32(5.99)
        s = np.float32(0)
        for _ in range(9):
            if not chi2 > th:
                s = np.float32(s + np.float32(th - chi2))
            if not chi1 > th:
                s = np.float32(s + np.float32(th - chi1))
        assert np.float32(o.score_H) == s, k


@pytest.mark.gpu
def test_cuda_geometry_validation_bit_exact(gpu_ctx, oracle):
    """pagk_geometry_validation through the C-ABI against the restatement (which the test above holds equal to the reference's
    own GeometryValidation()): status, both scores, the chosen model and the inlier count, bit for bit"""
    cases_g, cases_c = _cases(), _cases()
    big = make_case(21, n=3000, kind="translation", outliers=0.2, n_status0=300)     # more than one chunk of 1024
    cases_g.append(big); cases_c.append(make_case(21, n=3000, kind="translation", outliers=0.2, n_status0=300))
    gpu = []
    for k in range(0, len(cases_g), 8):
        gpu += gpu_ctx.geometry_validation(cases_g[k:k + 8])
    rc, cpu = oracle.geometry_validation(cases_c)
    assert rc == 0
    for k, (a, b, g, c) in enumerate(zip(cases_g, cases_c, gpu, cpu)):
        assert np.array_equal(a.out_status[:a.n_keys], b.out_status[:b.n_keys]), f"case {k}"
        assert (g.n_candidates, g.n_inlier, g.used_H) == (c.n_candidates, c.n_inlier, c.used_H), f"case {k}"
        for x, y in ((g.score_H, c.score_H), (g.score_F, c.score_F)):
            # a NaN score (singular model) is a NaN on both sides; the payload is the hardware's
            assert (np.isnan(x) and np.isnan(y)) or np.float32(x).view(np.uint32) == np.float32(y).view(np.uint32), f"case {k}: {x} != {y}"


@pytest.mark.gpu
def test_cuda_geometry_validation_on_resident_results(gpu_ctx, oracle):
    """TrackFeatures() then GeometryValidation() without moving the points: the second call reads the device-resident
    keys, predictions and status of the run before it"""
    pairs = [synth.make_pair(8700 + i, width=320, height=240, n_keys=300, pyramids=3, border=20) for i in range(3)]
    prm = capi.default_params(pyramids=3)
    outs = gpu_ctx.track_batch(pairs, prm)
    cases_g, cases_c = [], []
    for p, o in zip(pairs, outs):
        K = p.K.astype(np.float64)
        H = K @ o.Rcl.astype(np.float64) @ np.linalg.inv(K)           # the gyro homography stands in for findHomography's
        F = np.array([[0, -1e-3, 0.2], [1e-3, 0, -0.3], [-0.2, 0.3, 0.01]])
        g = capi.GeometryCase(None, None, None, H, F)
        g.resident_n_keys = p.n_keys
        cases_g.append(g)
        cases_c.append(capi.GeometryCase(p.keys_ref_un, o.pt_predict_un, o.status, H, F))
    gpu = gpu_ctx.geometry_validation(cases_g)
    rc, cpu = oracle.geometry_validation(cases_c)
    for a, b, g, c, p in zip(cases_g, cases_c, gpu, cpu, pairs):
        assert np.array_equal(a.out_status[:p.n_keys], b.out_status[:p.n_keys])
        assert (g.n_candidates, g.n_inlier, g.used_H) == (c.n_candidates, c.n_inlier, c.used_H)
        assert g.n_candidates > 8 and 0 < g.n_inlier <= g.n_candidates

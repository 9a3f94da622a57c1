"""CPU: analytic properties of the path, checked on the oracle (SURVEY.md section 4 / 8c golden list item 4)."""
import numpy as np

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth


def _pair(seed=31, **kw):
    args = dict(width=200, height=160, n_keys=60, pyramids=3, border=24, margin=32, K=synth.scaled_euroc_K(200))
    args.update(kw)
    return synth.make_pair(seed, **args)


def test_zero_rotation_predicts_identity(oracle):
    p = _pair()
    p.Rcl_override = np.eye(3, dtype=np.float32)
    o = oracle.gyro_predict(p, capi.default_params(e_type=1))
    assert np.allclose(o.pt_predict_un, p.keys_ref_un, atol=2e-4)
    assert o.status.all()
    assert np.allclose(o.affine.reshape(-1, 2, 2), np.eye(2), atol=1e-5)
    assert np.allclose(o.flows_predict_un, 0, atol=2e-4)


def test_roll_about_principal_axis_gives_rotation_matrix(oracle):
    p = _pair()
    th = 0.05
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    p.Rcl_override = R
    p.K = np.array([[300, 0, 100], [0, 300, 80], [0, 0, 1]], np.float32)   # fx == fy: K R K^-1 is a 2-D rotation about (cx, cy)
    o = oracle.gyro_predict(p, capi.default_params(e_type=1))
    ok = o.status.astype(bool)
    A = o.affine.reshape(-1, 2, 2)[ok]
    assert np.allclose(A, R[:2, :2], atol=2e-4)
    c = np.array([100, 80], np.float32)
    expect = (p.keys_ref_un - c) @ R[:2, :2].T + c
    assert np.allclose(o.pt_predict_un[ok], expect[ok], atol=2e-3)


def test_out_of_image_predictions_are_rejected(oracle):
    p = _pair(border=0, n_keys=300)
    th = 0.3
    p.Rcl_override = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)  # big yaw
    rc, o = oracle.track(p, capi.default_params(pyramids=3), 1)
    assert rc == 0 and (o.status == 0).any()
    bad = o.pt_gyro_predict_un[(o.pm_status == 0) & (o.iters == 0)]
    assert (bad == 0).all()          # skipped features keep the (0, 0) default


def test_flat_patch_fails_with_nan(oracle):
    """constant images: H is singular, update is NaN, the feature fails (src/patch_match.cpp:322-326)"""
    p = _pair()
    p.img_ref[:] = 128
    p.img_cur[:] = 128
    p.Rcl_override = np.eye(3, dtype=np.float32)
    rc, o = oracle.track(p, capi.default_params(pyramids=3), 1)
    assert rc == 0 and o.n_predict == 0 and (o.pm_status == 0).all() and (o.status == 0).all()


def test_integer_translation_is_recovered(oracle):
    """current = reference shifted by (3, -2): with no gyro (eType 5) LK must find the shift"""
    rng = np.random.default_rng(5)
    canvas = synth.texture(rng, 160, 200, margin=16, sigma=2.5)
    ref = np.clip(np.rint(canvas[16:176, 16:216]), 0, 255).astype(np.uint8)
    cur = np.clip(np.rint(canvas[18:178, 13:213]), 0, 255).astype(np.uint8)   # cur(x, y) = ref(x - 3, y + 2)
    keys = synth.random_keypoints(rng, 40, 200, 160, 30)
    p = capi.PairInputs(ref, cur, keys, np.array([0.0, 0.05]), np.zeros((2, 3), np.float32), 0.0, 0.05,
                        synth.scaled_euroc_K(200), np.eye(3))
    rc, o = oracle.track(p, capi.default_params(e_type=5, pyramids=3), 1)
    ok = o.status.astype(bool)
    assert rc == 0 and ok.sum() >= 30
    d = o.pt_predict_un[ok] - keys[ok]
    assert np.abs(d - np.array([3.0, -2.0])).max() < 0.05


def test_unsupported_types_return_minus_one(oracle):
    p = _pair()
    for e_type in (0, 7, -3):
        rc, o = oracle.track(p, capi.default_params(e_type=e_type), 1)
        assert rc == capi.PAGK_ERR_UNSUPPORTED and o.n_predict == -1


def test_get_pixel_value_clamps_and_wraps(oracle):
    img = (np.arange(12 * 9, dtype=np.uint8).reshape(9, 12) * 3) % 251
    f = np.float32
    assert oracle.get_pixel_value(img, -3.0, -2.0) == float(img[0, 0])
    assert oracle.get_pixel_value(img, 50.0, 50.0) == float(img[8, 11])       # clamped to the last pixel, weights 1,0
    x, y = f(11.5), f(2.0)                                                     # x in (cols-1, cols): tap wraps to next row
    expect = f(0.5) * f(img[2, 11]) + f(0.5) * f(img[3, 0])
    assert oracle.get_pixel_value(img, float(x), float(y)) == float(expect)
    x, y = f(4.0), f(8.25)                                                     # y in (rows-1, rows): guard row = last row
    assert oracle.get_pixel_value(img, float(x), float(y)) == float(img[8, 4])


def test_llt_solve_spd_and_breakdown(oracle):
    rng = np.random.default_rng(3)
    M = rng.normal(size=(4, 4))
    H = M @ M.T + 4 * np.eye(4)
    b = rng.normal(size=4)
    x = oracle.llt_solve(H, b)
    assert np.allclose(H @ x, b, atol=1e-10)
    Hz = np.zeros((4, 4))                       # first pivot 0: factorisation stops, solve divides 0/0 -> NaN
    assert np.isnan(oracle.llt_solve(Hz, np.zeros(4))[0])

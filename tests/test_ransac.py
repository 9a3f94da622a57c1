"""The device RANSAC of GeometryValidation (pagk_geometry_in.estimate = 1; csrc/pagk_ransac.h) -- statistical acceptance
against OpenCV's own estimators on 60 seeded fixtures (tests/golden/ransac_cv2.npz, made by tests/golden/make_ransac_golden.py
with cv2.findHomography(.., RANSAC, 3) and cv2.findFundamentalMat(.., FM_RANSAC, 3., 0.99), the two calls of reference
src/gyro_aided_tracker.cpp:597, :691).  Every comparison goes through the REFERENCE's scoring (CheckHomography /
CheckFundamental, chi-square 5.99 / 3.84, RH > 0.45) as restated in the oracle, once with OpenCV's models, once with ours:

  * pure rotation (the tracker's regime: a homography explains every correspondence): the same inlier set (IoU >= 0.99),
    the same model choice, the homography's score within 2 %
  * scenes with translation (general depth; a dominant plane plus clutter): our models are never worse than OpenCV's --
    at least 95 % of OpenCV's inliers are ours, at least 98 % as many inliers, the chosen model's score at least 98 % of
    OpenCV's.  (They are usually better: OpenCV's RANSAC stops at 99 % confidence on a plane-degenerate sample and refits
    the fundamental matrix on the plane's points; 1024 scored eight-point hypotheses do not fall for that.  An upper bound
    would reject the better model, so there is none.)

The CPU test runs the estimator code on the host (tests/cpp/ransac_host.cpp: the same header, sequentially); the GPU test
runs pagk_geometry_validation with estimate = 1 and also asks the device for the host harness's models.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_f32p, _u8p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("ransac") / "ransac_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", str(so),
                           os.path.join(ROOT, "tests", "cpp", "ransac_host.cpp")])
    return C.CDLL(str(so))


@pytest.fixture(scope="module")
def fixtures():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ransac_cv2.npz"))
    return [dict(p1=g[f"c{i}_p1"], p2=g[f"c{i}_p2"], status=g[f"c{i}_status"], H=g[f"c{i}_H"], F=g[f"c{i}_F"], kind=int(g[f"c{i}_kind"]))
            for i in range(int(g["n"]))]


def host_models(lib, c, seed, pair):
    H, F = np.zeros(9), np.zeros(9)
    rc = lib.pagk_ransac_host(len(c["status"]), c["p1"].ctypes.data_as(_f32p), c["p2"].ctypes.data_as(_f32p),
                              c["status"].ctypes.data_as(_u8p), seed, pair, H.ctypes.data_as(_f64p), F.ctypes.data_as(_f64p))
    assert rc == 0
    return H.reshape(3, 3), F.reshape(3, 3)


def scored(oracle, c, H, F):
    case = capi.GeometryCase(c["p1"], c["p2"], c["status"], H, F)
    rc, out = oracle.geometry_validation([case])
    assert rc == 0
    return case.out_status[:len(c["status"])].copy(), out[0]


def accept(c, mine, theirs):
    (a, oa), (b, ob) = mine, theirs
    inter, union = int(((a == 1) & (b == 1)).sum()), int(((a == 1) | (b == 1)).sum())
    if c["kind"] == 0:
        assert inter >= 0.99 * union, (inter, union)
        assert oa.used_H == ob.used_H == 1
        assert abs(oa.score_H - ob.score_H) <= 0.02 * ob.score_H
    assert inter >= 0.95 * int((b == 1).sum()), "inliers of OpenCV's model that ours rejects"
    assert oa.n_inlier >= 0.98 * ob.n_inlier
    chosen = lambda o: o.score_H if o.used_H else o.score_F
    assert chosen(oa) >= 0.98 * chosen(ob)
    assert oa.score_F >= 0.98 * ob.score_F


def test_host_estimators_against_opencv_fixtures(oracle, host_lib, fixtures):
    assert len(fixtures) >= 50
    for i, c in enumerate(fixtures):
        H, F = host_models(host_lib, c, 1, i)
        assert abs(H[2, 2] - 1.0) < 1e-12 and np.isfinite(H).all() and np.isfinite(F).all()
        assert abs(np.linalg.det(F)) <= 1e-10 * np.abs(F).max() ** 3 + 1e-18      # rank two
        accept(c, scored(oracle, c, H, F), scored(oracle, c, c["H"], c["F"]))


def test_host_estimators_do_not_depend_much_on_the_seed(oracle, host_lib, fixtures):
    for i, c in list(enumerate(fixtures))[::6]:
        ref = scored(oracle, c, *host_models(host_lib, c, 1, i))
        for seed in (2, 3):
            a = scored(oracle, c, *host_models(host_lib, c, seed, i))
            inter, union = int(((a[0] == 1) & (ref[0] == 1)).sum()), int(((a[0] == 1) | (ref[0] == 1)).sum())
            assert inter >= 0.93 * union and a[1].used_H == ref[1].used_H, (i, seed, inter, union)


@pytest.mark.gpu
def test_device_estimators(gpu_ctx, oracle, host_lib, fixtures):
    """pagk_geometry_validation with estimate = 1: the device's models are the host harness's (same code, same generator),
    its scoring is the oracle's on those models, and the acceptance against OpenCV holds"""
    batch = 8
    for b0 in range(0, len(fixtures), batch):
        group = fixtures[b0:b0 + batch]
        cases = [capi.GeometryCase(c["p1"], c["p2"], c["status"], estimate=True, seed=1) for c in group]
        outs = gpu_ctx.geometry_validation(cases)
        for k, (c, case, o) in enumerate(zip(group, cases, outs)):
            Hd, Fd = np.array(o.H21).reshape(3, 3), np.array(o.F21).reshape(3, 3)
            Hh, Fh = host_models(host_lib, c, 1, k)      # the kernel's `pair` is the index inside the call
            assert np.allclose(Hd, Hh, rtol=1e-6, atol=1e-9), (b0 + k, Hd, Hh)
            assert np.allclose(Fd, Fh, rtol=1e-5, atol=1e-9), (b0 + k, Fd, Fh)
            st, oo = scored(oracle, c, Hd, Fd)
            assert np.array_equal(case.out_status[:len(st)], st)
            assert (o.used_H, o.n_inlier, o.n_candidates) == (oo.used_H, oo.n_inlier, oo.n_candidates)
            assert o.score_H == oo.score_H and o.score_F == oo.score_F
            accept(c, (st, oo), scored(oracle, c, c["H"], c["F"]))


@pytest.mark.gpu
def test_device_estimators_on_resident_results(gpu_ctx, oracle):
    """the drivers' sequence TrackFeatures() -> GeometryValidation() with nothing leaving the device in between"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import synth
    pairs = [synth.make_pair(9400 + i, width=752, height=480, n_keys=600, pyramids=3) for i in range(3)]
    outs = gpu_ctx.track_batch(pairs, capi.default_params(pyramids=3))
    cases = [capi.GeometryCase(None, None, None, estimate=True, seed=7) for _ in pairs]
    for c, p in zip(cases, pairs):
        c.resident_n_keys = p.n_keys
    res = gpu_ctx.geometry_validation(cases)
    for p, o, r, c in zip(pairs, outs, res, cases):
        assert r.n_candidates == int(o.status.sum()) and r.used_H == 1       # a rotating camera: the homography wins
        assert r.n_inlier >= 0.9 * r.n_candidates
        st = c.out_status[:p.n_keys]
        assert int((st == 1).sum()) == r.n_inlier and not ((st == 1) & (o.status == 0)).any()

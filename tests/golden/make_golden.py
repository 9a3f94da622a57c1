"""Generates tests/golden/*.npz.  Run HERE (needs cv2 4.x; the GPU box and the tests never run it).

What is pinned against what (the reference ships no tests or golden vectors):
  pyramid.npz   cv2.resize(INTER_LINEAR) chains  -> pins oracle resize_half() and the CUDA pyramid kernel
  matexpr.npz   cv2.gemm / cv2.invert / cv2.scaleAdd / cv2.add chains that OpenCV's cv::MatExpr lowers
                IntegrateGyroMeasurements, SetRcl and the affine-matrix expression to
                (src/gyro_aided_tracker.cpp:166-167, 511-587)       -> pins oracle small_*() and integrate_gyro()
  remap.npz     cv2.remap (INTER_LINEAR, u8, float maps, constant border)               -> pins oracle remap_linear()
  fast.npz      cv2.FastFeatureDetector (TYPE_9_16) keypoints and responses            -> pins oracle fast_detect()
  lk_frozen.npz outputs of the reference build (oracle/_ref/libpagk_ref.so = the reference's own three sources
                compiled against stand-in OpenCV/Eigen/glog headers, oracle/reference.py) on small seeded pairs
                -> pins the restatement's Gauss-Newton loop, prediction, filter and control flow against the
                reference's own code.  Eigen's LLT/norm arithmetic inside it is still a restatement (Eigen is not
                available): "parity unpinned" for that one piece.
"""
import ctypes
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

f32 = np.float32
libm = ctypes.CDLL("libm.so.6")
libm.sinf.restype = libm.cosf.restype = ctypes.c_float
libm.sinf.argtypes = libm.cosf.argtypes = [ctypes.c_float]


def lcg_image(h, w, seed):
    x = (np.arange(h * w, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(seed * 1442695040888963407 + 1))
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    return ((x >> np.uint64(40)) & np.uint64(0xFF)).astype(np.uint8).reshape(h, w)


def make_pyramid():
    out = {}
    shapes = [(480, 752, 5), (480, 640, 4), (1080, 1920, 5), (2160, 3840, 5), (135, 240, 4), (77, 101, 4), (270, 135, 3),
              (64, 64, 5), (67, 135, 3), (33, 50, 3)]
    for (h, w, levels) in shapes:
        img = lcg_image(h, w, 7)
        cur = img
        for l in range(1, levels):
            dsz = (int(cur.shape[1] * 0.5), int(cur.shape[0] * 0.5))
            if dsz[0] < 1 or dsz[1] < 1:
                break
            cur = cv2.resize(cur, dsz)   # default INTER_LINEAR, as src/patch_match.cpp:69
            key = f"{h}x{w}_L{l}"
            out["sha_" + key] = np.frombuffer(hashlib.sha256(cur.tobytes()).digest(), np.uint8)
            out["shape_" + key] = np.array(cur.shape, np.int32)
            if h * w <= 135 * 240:
                out["img_" + key] = cur
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **out)
    print("pyramid.npz", len(out))


def integrate_cv2(imu_t, imu_w, t_ref, t_cur, bias, Rbc, K):
    """IntegrateGyroMeasurements + SetRcl with cv2 primitives standing in for cv::MatExpr"""
    I = np.eye(3, dtype=f32)
    dR = I.copy()
    n = len(imu_t) - 1
    for i in range(n):
        w0, w1 = imu_w[i].astype(f32), imu_w[i + 1].astype(f32)
        if i == 0 and i < n - 1:
            tab = f32(imu_t[i + 1] - imu_t[i]); tini = f32(imu_t[i] - t_ref); r = f32(tini / tab)
            av = ((w0 + w1) - (w1 - w0) * r) * f32(0.5); tstep = f32(imu_t[i + 1] - t_ref)
        elif i < n - 1:
            av = (w0 + w1) * f32(0.5); tstep = f32(imu_t[i + 1] - imu_t[i])
        elif i > 0 and i == n - 1:
            tab = f32(imu_t[i + 1] - imu_t[i]); tend = f32(imu_t[i + 1] - t_cur); r = f32(tend / tab)
            av = ((w0 + w1) - (w1 - w0) * r) * f32(0.5); tstep = f32(t_cur - imu_t[i])
        else:
            av = w0; tstep = f32(t_cur - t_ref)
        x, y, z = [f32(float(f32(av[k] - bias[k])) * float(tstep)) for k in range(3)]
        d2 = f32(f32(f32(x * x) + f32(y * y)) + f32(z * z))
        d = f32(np.sqrt(d2))
        W = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], f32)
        if float(d) < 1e-4:
            dRi = cv2.add(I, W)
        else:
            a = float(libm.sinf(float(d))) * (1.0 / float(d))
            b = float(f32(f32(1.0) - f32(libm.cosf(float(d))))) * (1.0 / float(d2))
            dRi = cv2.add(cv2.scaleAdd(W, a, I), cv2.gemm(W, W, b, None, 0))
        dR = cv2.gemm(dR, dRi, 1, None, 0)
    M1 = cv2.gemm(Rbc, dR, 1, None, 0, flags=cv2.GEMM_1_T | cv2.GEMM_2_T)
    Rcl = cv2.gemm(M1, Rbc, 1, None, 0)
    return Rcl, krk_cv2(K, Rcl)


def krk_cv2(K, Rcl):
    return cv2.gemm(cv2.gemm(K, Rcl, 1, None, 0), cv2.invert(K)[1], 1, None, 0)


def affine_cv2(cflows, half):
    C = np.ascontiguousarray(cflows.reshape(4, 2).T)                      # matC 2x4
    B = np.array([[-half, half, -half, half], [-half, -half, half, half]], f32)
    S = cv2.gemm(C, B, 1, None, 0, flags=cv2.GEMM_2_T)
    BBt = cv2.gemm(B, B, 1, None, 0, flags=cv2.GEMM_2_T)
    return cv2.gemm(S, cv2.invert(BBt)[1], 1, None, 0)


def make_matexpr():
    rng = np.random.default_rng(20261018)
    out = {"n": np.array(48)}
    for i in range(48):
        K = synth.EUROC_K.copy()
        K[0, 0] += f32(rng.uniform(-40, 40)); K[1, 1] += f32(rng.uniform(-40, 40))
        K[0, 2] += f32(rng.uniform(-20, 20)); K[1, 2] += f32(rng.uniform(-20, 20))
        Rbc = synth.EUROC_RBC if i % 3 else np.eye(3, dtype=f32)
        n_imu = [2, 3, 5, 12, 17][i % 5]
        t_ref = 1403715000.0 + i
        imu_t = t_ref - rng.uniform(0, 0.004) + np.arange(n_imu) * 0.005
        t_cur = imu_t[-1] - rng.uniform(0, 0.004) if n_imu > 2 else imu_t[-1] + 0.001
        scale = [1e-6, 0.05, 0.5, 3.0][i % 4]           # includes the d < 1e-4 branch
        imu_w = (rng.normal(0, scale, (n_imu, 3))).astype(f32)
        bias = rng.normal(0, 0.01, 3).astype(f32) if i % 2 else np.zeros(3, f32)
        Rcl, KRK = integrate_cv2(imu_t, imu_w, t_ref, t_cur, bias, Rbc, K)
        for k, v in dict(K=K, Rbc=Rbc, imu_t=imu_t, imu_w=imu_w, t_ref=np.array(t_ref), t_cur=np.array(t_cur), bias=bias,
                         Rcl=Rcl, KRK=KRK).items():
            out[f"g{i}_{k}"] = np.asarray(v)
        half = [3, 5, 7, 10][i % 4]
        cfl = (np.array([[-half, -half], [half, -half], [-half, half], [half, half]], f32) *
               f32(1 + rng.normal(0, 0.05)) + rng.normal(0, 0.3, (4, 2)).astype(f32)).astype(f32)
        out[f"a{i}_half"] = np.array(half); out[f"a{i}_cflows"] = cfl; out[f"a{i}_A"] = affine_cv2(cfl, half)
    np.savez_compressed(os.path.join(HERE, "matexpr.npz"), **out)
    print("matexpr.npz", len(out))


def make_lk_frozen():
    """outputs of the REFERENCE BUILD (oracle/_ref/libpagk_ref.so: the reference's own sources compiled against the
    stand-in headers, oracle/reference.py) on small seeded pairs; the restatement must agree bit for bit before the
    fixture is written.  `iters` (per-feature pass counts) is not observable from the reference and comes from the
    restatement; its sum is checked against the reference build's count of H.llt() calls."""
    from oracle import reference
    from tests import helpers
    reference.build(force=True)
    out = {}
    cases = [dict(seed=9001, e_type=4, dist=None), dict(seed=9002, e_type=3, dist=None),
             dict(seed=9003, e_type=4, dist=synth.EUROC_DIST), dict(seed=9004, e_type=2, dist=None),
             dict(seed=9005, e_type=5, dist=None), dict(seed=9006, e_type=6, dist=None),
             # TrackFeatures() hard-codes 3 levels / 10 iterations; these go through the composed path
             dict(seed=9007, e_type=4, dist=None, pyramids=4), dict(seed=9008, e_type=4, dist=None, half_patch=10, pyramids=2),
             dict(seed=9009, e_type=4, dist=None, border=0, sigma_w=3.0), dict(seed=9010, e_type=4, dist=None, flat=True)]
    out["n"] = np.array(len(cases))
    for i, c in enumerate(cases):
        pyr, half = c.get("pyramids", 3), c.get("half_patch", 5)
        p = synth.make_pair(c["seed"], width=160, height=120, n_keys=48, pyramids=pyr, half_patch=half,
                            border=c.get("border", 12 if half == 5 else 24), margin=32,
                            K=synth.scaled_euroc_K(160), dist=c["dist"], sigma_w=c.get("sigma_w", 0.8))
        if c.get("flat"):   # saturated / constant regions: singular normal matrix, NaN update, status 0
            p.img_ref[:, :80] = 255
            p.img_cur[:, :80] = 255
            p.img_ref[60:, 80:] = 0
            p.img_cur[60:, 80:] = 0
        prm = capi.default_params(e_type=c["e_type"], pyramids=pyr, half_patch=half)
        rc, o = reference.track(p, prm, 1)
        assert rc == 0
        rc2, o2 = oracle.track(p, prm, 1)
        assert rc2 == 0
        rep = helpers.compare(o, o2)
        bad = {k: v for k, v in rep.items() if isinstance(v, dict) and v.get("bit_mismatch", 0) and k != "iters"}
        assert not bad and o.n_predict == o2.n_predict and o.n_iterations == o2.n_iterations, (i, bad)
        assert rep["Rcl_bits"] == 0 and rep["KRKinv_bits"] == 0
        for k, v in dict(img_ref=p.img_ref, img_cur=p.img_cur, keys=p.keys_ref_un, imu_t=p.imu_t, imu_w=p.imu_w,
                         t_ref=np.array(p.t_ref), t_cur=np.array(p.t_cur), K=p.K, Rbc=p.Rbc, dist=p.dist,
                         n_dist=np.array(p.n_dist), e_type=np.array(c["e_type"]), pyramids=np.array(pyr),
                         half_patch=np.array(half)).items():
            out[f"c{i}_in_{k}"] = np.asarray(v)
        for k, v in o.arrays().items():
            out[f"c{i}_out_{k}"] = v if k != "iters" else o2.iters
        out[f"c{i}_out_Rcl"] = o.Rcl; out[f"c{i}_out_KRKinv"] = o.KRKinv
        out[f"c{i}_out_n_predict"] = np.array(o.n_predict); out[f"c{i}_out_n_iterations"] = np.array(o.n_iterations)
        out[f"c{i}_ref_path"] = np.array(reference.last_path())
        print("case", i, "eType", c["e_type"], "n_predict", o.n_predict, "iters", o.n_iterations, "path", reference.last_path(),
              "status0", int((o.status == 0).sum()), "nan", int(np.isnan(o.pixel_error).sum()))
    np.savez_compressed(os.path.join(HERE, "lk_frozen.npz"), **out)


def make_fast():
    """cv2.FastFeatureDetector (TYPE_9_16) keypoints: positions in OpenCV's order and responses -> pins oracle fast_detect()"""
    rng = np.random.default_rng(77)
    out, cases = {}, [(60, 80, 10, 1.2), (240, 320, 20, 2.0), (480, 752, 7, 2.0), (100, 101, 0, 0.0), (37, 50, 40, 0.8), (480, 640, 20, 1.5)]
    out["n"] = np.array(len(cases))
    for i, (h, w, th, sig) in enumerate(cases):
        img = (rng.random((h, w)) * 255).astype(np.uint8)
        if sig > 0:
            img = cv2.normalize(cv2.GaussianBlur(img, (0, 0), sig), None, 0, 255, cv2.NORM_MINMAX)
        out[f"f{i}_seed_shape_th"] = np.array([h, w, th], np.int32)
        if h * w <= 240 * 320:
            out[f"f{i}_img"] = img
        out[f"f{i}_img_sha"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
        for nm in (1, 0):
            k = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=bool(nm), type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(img)
            xy = np.array([p.pt for p in k], np.float32).reshape(-1, 2)
            rs = np.array([p.response for p in k], np.float32)
            out[f"f{i}_n{nm}"] = np.array(len(k))
            out[f"f{i}_sha{nm}"] = np.frombuffer(hashlib.sha256(xy.tobytes() + rs.tobytes()).digest(), np.uint8)
            if h * w <= 240 * 320 and nm:
                out[f"f{i}_xy"] = xy; out[f"f{i}_rs"] = rs
    np.savez_compressed(os.path.join(HERE, "fast.npz"), **out)
    print("fast.npz", len(out))


def make_remap():
    """cv2.remap(INTER_LINEAR) on u8 with float maps -> pins oracle remap_linear(): random maps with out-of-range, integer and tie
    coordinates, and the rectification map of the EuRoC calibration (cv2.initUndistortRectifyMap, as include/imu_types.h:63-65)"""
    rng = np.random.default_rng(99)
    out = {}
    H, W = 60, 80
    img = (rng.random((H, W)) * 255).astype(np.uint8)
    mx = rng.uniform(-5, W + 4, (48, 64)).astype(np.float32); my = rng.uniform(-5, H + 4, (48, 64)).astype(np.float32)
    mx[::7, ::5] = np.round(mx[::7, ::5]); my[::3, ::11] = np.round(my[::3, ::11])
    mx[1, :10] = [-1, -0.99, -1.01, W - 1, W - 1.01, W - 0.5, W, 1e6, -1e6, 40000.0]; my[1, :10] = 0
    my[2, :6] = [-1, -0.99, H - 1, H - 1.01, H, 1e6]; mx[2, :6] = 3
    mx[3, :] = (np.arange(64) + 1 / 64).astype(np.float32); my[3, :] = np.float32(7 + 3 / 64)     # ties of the 1/32 grid
    out["s_img"], out["s_mx"], out["s_my"] = img, mx, my
    out["s_out"] = cv2.remap(img, mx, my, cv2.INTER_LINEAR)
    H, W = 480, 752
    K = synth.EUROC_K.astype(np.float64); D = np.asarray(synth.EUROC_DIST, np.float64)
    newK = cv2.getOptimalNewCameraMatrix(K, D, (W, H), 0, (W, H))[0]
    M1, M2 = cv2.initUndistortRectifyMap(K, D, None, newK, (W, H), cv2.CV_32F)
    big = lcg_image(H, W, 11)
    res = cv2.remap(big, M1, M2, cv2.INTER_LINEAR)
    out["r_M1"], out["r_M2"] = M1.astype(np.float16).astype(np.float32), M2.astype(np.float16).astype(np.float32)   # halves keep the fixture small
    res = cv2.remap(big, out["r_M1"], out["r_M2"], cv2.INTER_LINEAR)
    out["r_sha"] = np.frombuffer(hashlib.sha256(res.tobytes()).digest(), np.uint8)
    out["r_M1"], out["r_M2"] = out["r_M1"].astype(np.float16), out["r_M2"].astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "remap.npz"), **out)
    print("remap.npz", len(out))


if __name__ == "__main__":
    make_pyramid()
    make_matexpr()
    make_lk_frozen()
    make_fast()
    make_remap()

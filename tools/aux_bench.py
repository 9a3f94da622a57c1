"""Times the kernels around the hot path at BASELINE config B's shapes, through the C-ABI:
wall clock per call (host buffers in and out), and -- run under `ncu --metrics gpu__time_duration.sum` -- device time per kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker

W, H, N, P = 752, 480, 1024, 64
cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
pairs = [synth.make_pair(2000 + i, **cfg) for i in range(8)]
pairs = [pairs[i % 8] for i in range(P)]
prm = capi.default_params(pyramids=4)
reps = int(os.environ.get("REPS", "20"))


def timed(name, fn, unit, per_call):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:46s} {1e3 * dt:8.3f} ms per call   {per_call / dt / 1e6:9.2f} M {unit}/s")


with tracker.Context(max_width=W, max_height=H, max_keys=N, max_pairs=P, max_levels=4) as ctx:
    outs = ctx.track_batch(pairs, prm)
    K = pairs[0].K
    Hm = [K.astype(np.float64) @ o.Rcl.astype(np.float64) @ np.linalg.inv(K.astype(np.float64)) for o in outs]
    F = np.array([[0, -1e-3, 0.2], [1e-3, 0, -0.3], [-0.2, 0.3, 0.01]])
    geo = []
    for p, h in zip(pairs, Hm):
        g = capi.GeometryCase(None, None, None, h, F)
        g.resident_n_keys = p.n_keys
        geo.append(g)
    timed("GeometryValidation scoring, 64 pairs resident", lambda: ctx.geometry_validation(geo), "features", P * N)
    ln = ((pairs[0].keys_ref_un - [K[0, 2], K[1, 2]]) / [K[0, 0], K[1, 1]]).astype(np.float32)
    carry = [capi.CarryCase(None, None, None, ln, K, p.t_cur, p.t_ref, W, H, n_keys=p.n_keys) for p in pairs]
    timed("SetPredictKeyPointsAndMask, 64 pairs + masks", lambda: ctx.set_predict_keypoints_and_mask(carry), "features", P * N)
    carry_nm = [capi.CarryCase(None, None, None, ln, K, p.t_cur, p.t_ref, W, H, n_keys=p.n_keys, want_mask=False) for p in pairs]
    timed("SetPredictKeyPointsAndMask, 64 pairs, no mask D2H", lambda: ctx.set_predict_keypoints_and_mask(carry_nm), "features", P * N)
    img = pairs[0].img_cur
    timed("cv::FAST 752x480, one image", lambda: ctx.fast_detect(img, 20, True, max_out=20000), "pixels", W * H)
    timed("per-cell FAST (ORB top-up) 752x480, one image", lambda: ctx.orb_cell_detect(img, 20, 7, mask=carry[0].mask, max_out=20000), "pixels", W * H)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    mx, my = (xs + 1.5 * np.sin(ys / 37.0)).astype(np.float32), (ys + 1.2 * np.cos(xs / 41.0)).astype(np.float32)
    timed("cv::remap 752x480, one image (maps uploaded)", lambda: ctx.remap_linear(img, mx, my), "pixels", W * H)
    ctx.set_rectify_maps(mx, my)
    timed("track 64 pairs with in-pipeline rectification", lambda: ctx.track_batch(pairs, prm), "features", P * N)
    ctx.set_rectify_maps(None, None)
    timed("track 64 pairs (pageable host buffers)", lambda: ctx.track_batch(pairs, prm), "features", P * N)

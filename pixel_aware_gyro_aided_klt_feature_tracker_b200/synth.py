"""Seeded synthetic frame pairs + gyro for the hot path.

Texture = Gaussian-blurred (sigma 2 px) uniform noise on a padded canvas; the current frame is the
K.R.K^-1 warp of the reference frame, with gain U(0.9,1.1), bias U(-8,8) and sigma=1 grey-level noise;
gyro at 200 Hz with a timestamp offset so that the first/last-interval interpolation branches of
IntegrateGyroMeasurements (reference src/gyro_aided_tracker.cpp:530-555) run.  numpy only: no cv2, no
network, no dataset.  Calibrations are the reference's own yaml files.
"""
from __future__ import annotations

import numpy as np

from .capi import PairInputs

# Examples/ROS/ROS_Demo_Feature_Tracking/config/EuRoC.yaml:33-58
EUROC_K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
EUROC_DIST = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
EUROC_RBC = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422],
                      [0.999557249008, 0.0149672133247, 0.025715529948],
                      [-0.0257744366974, 0.00375618835797, 0.999660727178]], np.float32)
# Examples/Demo/RealSenseD435i.yaml:27-51
D435I_K = np.array([[394.5643528049837, 0, 325.2710790421636], [0, 395.2103902700227, 243.20141864231425],
                    [0, 0, 1]], np.float32)
D435I_DIST = np.array([-0.0027697209770466296, -0.0007212258451583873, 0.00029903960869777114,
                       0.0003981049435158156], np.float32)
D435I_RBC = np.eye(3, dtype=np.float32)

#: the BASELINE.json configs (A..E of SURVEY.md section 8a)
CONFIGS = {
    "A": dict(width=640, height=480, n_keys=500, half_patch=5, pyramids=3, fps=15, sigma_w=0.5, K=D435I_K,
              Rbc=D435I_RBC, pairs=1),
    "B": dict(width=752, height=480, n_keys=1024, half_patch=5, pyramids=4, fps=20, sigma_w=0.5, K=EUROC_K,
              Rbc=EUROC_RBC, pairs=64),
    "C": dict(width=1920, height=1080, n_keys=8192, half_patch=10, pyramids=4, fps=20, sigma_w=0.5,
              K=None, Rbc=EUROC_RBC, pairs=1),
    "D": dict(width=752, height=480, n_keys=1024, half_patch=5, pyramids=4, fps=20, sigma_w=0.5, K=EUROC_K,
              Rbc=EUROC_RBC, pairs=256),
    "E": dict(width=3840, height=2160, n_keys=32768, half_patch=5, pyramids=5, fps=20, sigma_w=3.0,
              K=None, Rbc=EUROC_RBC, pairs=1),
}


def scaled_euroc_K(width: int) -> np.ndarray:
    s = width / 752.0
    K = EUROC_K.astype(np.float64).copy()
    K[:2] *= s
    return K.astype(np.float32)


def _gauss_blur(a: np.ndarray, sigma: float) -> np.ndarray:
    r = int(np.ceil(3 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    try:
        from scipy.ndimage import correlate1d
        a = correlate1d(a, k, axis=0, mode="reflect")
        return correlate1d(a, k, axis=1, mode="reflect")
    except Exception:  # pragma: no cover - scipy is in the image
        a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, "valid"), 0, a)
        return np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, "valid"), 1, a)


def texture(rng: np.random.Generator, height: int, width: int, margin: int = 64, sigma: float = 2.0) -> np.ndarray:
    """float64 canvas in [0,255], (height+2m) x (width+2m)"""
    t = _gauss_blur(rng.random((height + 2 * margin, width + 2 * margin)), sigma)
    t -= t.min()
    t *= 255.0 / t.max()
    return t


def so3_exp(v: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(v))
    W = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], np.float64)
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + W * (np.sin(th) / th) + W @ W * ((1 - np.cos(th)) / th ** 2)


def _bilinear(canvas: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    h, w = canvas.shape
    x = np.clip(x, 0, w - 1.001)
    y = np.clip(y, 0, h - 1.001)
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    fx, fy = x - x0, y - y0
    return ((1 - fy) * ((1 - fx) * canvas[y0, x0] + fx * canvas[y0, x0 + 1]) +
            fy * ((1 - fx) * canvas[y0 + 1, x0] + fx * canvas[y0 + 1, x0 + 1]))


def random_keypoints(rng: np.random.Generator, n: int, width: int, height: int, border: int) -> np.ndarray:
    """n points >= border px from every edge, half integer-valued (fresh detections) and half sub-pixel
    (carried-over tracks), de-duplicated on a 3 px grid."""
    out = np.zeros((0, 2), np.float64)
    seen = set()
    while out.shape[0] < n:
        m = 2 * (n - out.shape[0]) + 16
        p = np.stack([rng.uniform(border, width - border, m), rng.uniform(border, height - border, m)], 1)
        keep = []
        for q in p:
            c = (int(q[0] // 3), int(q[1] // 3))
            if c not in seen:
                seen.add(c)
                keep.append(q)
        out = np.concatenate([out, np.array(keep).reshape(-1, 2)])[:n]
    out[: n // 2] = np.round(out[: n // 2])
    return out.astype(np.float32)


def make_pair(seed: int, width: int = 752, height: int = 480, n_keys: int = 1024, half_patch: int = 5,
              pyramids: int = 4, fps: float = 20.0, sigma_w: float = 0.5, K=None, Rbc=None, dist=None,
              imu_rate: float = 200.0, border: int | None = None, margin: int = 64, **_unused) -> PairInputs:
    rng = np.random.default_rng(seed)
    K = scaled_euroc_K(width) if K is None else np.asarray(K, np.float32)
    Rbc = EUROC_RBC if Rbc is None else np.asarray(Rbc, np.float32)
    K64, Rbc64 = K.astype(np.float64), Rbc.astype(np.float64)
    canvas = texture(rng, height, width, margin)
    ref = np.clip(np.rint(canvas[margin:margin + height, margin:margin + width]), 0, 255).astype(np.uint8)

    dt = 1.0 / fps
    w_body = rng.normal(0.0, sigma_w, 3)
    # tracker: Rcl = Rbc^T . dR^T . Rbc with dR = exp(w_body dt)  (src/gyro_aided_tracker.cpp:560)
    Rcl = Rbc64.T @ so3_exp(w_body * dt).T @ Rbc64
    Hcl = K64 @ Rcl @ np.linalg.inv(K64)          # p_cur ~ Hcl p_ref
    Hlc = np.linalg.inv(Hcl)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float64)
    den = Hlc[2, 0] * xs + Hlc[2, 1] * ys + Hlc[2, 2]
    sx = (Hlc[0, 0] * xs + Hlc[0, 1] * ys + Hlc[0, 2]) / den + margin
    sy = (Hlc[1, 0] * xs + Hlc[1, 1] * ys + Hlc[1, 2]) / den + margin
    cur = _bilinear(canvas, sx, sy)
    gain, bias = rng.uniform(0.9, 1.1), rng.uniform(-8.0, 8.0)
    cur = cur * gain + bias + rng.normal(0.0, 1.0, cur.shape)
    cur = np.clip(np.rint(cur), 0, 255).astype(np.uint8)

    t_ref = 1403715000.0 + seed * 0.05     # EuRoC-like absolute timestamps (seconds)
    t_cur = t_ref + dt
    off = rng.uniform(0.0, 0.005)
    n_imu = int(np.ceil(dt * imu_rate)) + 2
    imu_t = t_ref - off + np.arange(n_imu) / imu_rate
    # keep the last sample at or after t_cur so the last-interval branch interpolates
    while imu_t[-1] < t_cur:
        imu_t = np.append(imu_t, imu_t[-1] + 1.0 / imu_rate)
    imu_w = w_body[None, :] + rng.normal(0.0, 1.7e-4 * np.sqrt(imu_rate), (imu_t.size, 3))

    if border is None:
        border = 8 * (half_patch + 3) if pyramids >= 4 else 4 * (half_patch + 3)
        border = min(border, min(width, height) // 4)
    keys = random_keypoints(rng, n_keys, width, height, border)
    d = np.zeros(4, np.float32) if dist is None else np.asarray(dist, np.float32)
    return PairInputs(ref, cur, keys, imu_t, imu_w.astype(np.float32), t_ref, t_cur, K, Rbc, dist=d, n_dist=d.size)


def make_sequence(seed: int, n_frames: int, width: int = 752, height: int = 480, n_keys: int = 1024, half_patch: int = 5,
                  pyramids: int = 4, fps: float = 20.0, sigma_w: float = 0.5, K=None, Rbc=None, imu_rate: float = 200.0,
                  border: int | None = None, margin: int = 64, **_unused):
    """One camera + IMU stream: frame 0 and n_frames - 1 consecutive pairs (frame t-1 -> frame t) of one rotating camera over
    one texture.  Returns (frames, pairs): pairs[t-1] is the PairInputs of (frames[t-1], frames[t]) with fresh keypoints on
    frames[t-1], as a driver that re-detects every frame would hand them over."""
    rng = np.random.default_rng(seed)
    K = scaled_euroc_K(width) if K is None else np.asarray(K, np.float32)
    Rbc = EUROC_RBC if Rbc is None else np.asarray(Rbc, np.float32)
    K64, Rbc64 = K.astype(np.float64), Rbc.astype(np.float64)
    Kinv = np.linalg.inv(K64)
    canvas = texture(rng, height, width, margin)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float64)
    dt = 1.0 / fps
    if border is None:
        border = 8 * (half_patch + 3) if pyramids >= 4 else 4 * (half_patch + 3)
        border = min(border, min(width, height) // 4)
    R_acc = np.eye(3)                       # camera rotation of frame t relative to frame 0: p_t ~ K R_acc K^-1 p_0
    frames, pairs = [], []
    t0 = 1403715000.0 + seed * 0.05
    for t in range(n_frames):
        Hl0 = np.linalg.inv(K64 @ R_acc @ Kinv)
        den = Hl0[2, 0] * xs + Hl0[2, 1] * ys + Hl0[2, 2]
        sx = (Hl0[0, 0] * xs + Hl0[0, 1] * ys + Hl0[0, 2]) / den + margin
        sy = (Hl0[1, 0] * xs + Hl0[1, 1] * ys + Hl0[1, 2]) / den + margin
        img = _bilinear(canvas, sx, sy) * rng.uniform(0.95, 1.05) + rng.uniform(-4.0, 4.0) + rng.normal(0.0, 1.0, (height, width))
        frames.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
        if t + 1 == n_frames:
            break
        w_body = rng.normal(0.0, sigma_w, 3)
        Rcl = Rbc64.T @ so3_exp(w_body * dt).T @ Rbc64
        t_ref, t_cur = t0 + t * dt, t0 + (t + 1) * dt
        off = rng.uniform(0.0, 0.005)
        imu_t = t_ref - off + np.arange(int(np.ceil(dt * imu_rate)) + 2) / imu_rate
        while imu_t[-1] < t_cur:
            imu_t = np.append(imu_t, imu_t[-1] + 1.0 / imu_rate)
        imu_w = w_body[None, :] + rng.normal(0.0, 1.7e-4 * np.sqrt(imu_rate), (imu_t.size, 3))
        keys = random_keypoints(rng, n_keys, width, height, border)
        pairs.append(dict(keys=keys, imu_t=imu_t, imu_w=imu_w.astype(np.float32), t_ref=t_ref, t_cur=t_cur))
        R_acc = Rcl @ R_acc
    out = [PairInputs(frames[k], frames[k + 1], q["keys"], q["imu_t"], q["imu_w"], q["t_ref"], q["t_cur"], K, Rbc)
           for k, q in enumerate(pairs)]
    return frames, out


def make_config_pairs(name: str, n_pairs: int | None = None, seed0: int | None = None, **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    n = cfg.pop("pairs") if n_pairs is None else (cfg.pop("pairs"), n_pairs)[1]
    base = 1000 * (ord(name) - ord("A") + 1) if seed0 is None else seed0
    return [make_pair(base + i, **cfg) for i in range(n)], cfg

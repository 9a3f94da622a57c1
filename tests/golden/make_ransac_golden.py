"""Fixtures for the statistical acceptance of the device RANSAC (tests/test_ransac.py): seeded correspondence sets and what
OpenCV's own estimators return for them -- cv2.findHomography(p1, p2, cv2.RANSAC, 3) and
cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 3., 0.99), the two calls of GyroAidedTracker::GeometryValidation
(reference src/gyro_aided_tracker.cpp:597, :691).  Needs cv2 (the build container has 4.13); run from the repository root:

    python tests/golden/make_ransac_golden.py        ->  tests/golden/ransac_cv2.npz
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float64)   # EuRoC cam0


def rodrigues(w):
    R, _ = cv2.Rodrigues(np.asarray(w, np.float64))
    return R


def make_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(60, 1024))
    p1 = np.stack([rng.uniform(20, 732, n), rng.uniform(20, 460, n)], axis=1)
    kind = seed % 3
    R = rodrigues(rng.normal(0, 0.03, 3))
    ray = np.linalg.inv(K) @ np.concatenate([p1, np.ones((n, 1))], axis=1).T          # 3 x n
    if kind == 0:      # pure rotation: the tracker's usual case, a homography explains everything
        X = ray * rng.uniform(2, 20, n)
        t = np.zeros(3)
    elif kind == 1:    # a general scene with translation: only the fundamental matrix explains it
        X = ray * rng.uniform(2, 12, n)
        t = rng.normal(0, 0.15, 3)
    else:              # a dominant plane plus clutter, with translation
        depth = 5.0 / np.maximum(0.2, (np.array([0.1, 0.05, 1.0]) @ ray))
        clutter = rng.random(n) < 0.3
        depth[clutter] = rng.uniform(2, 12, clutter.sum())
        X = ray * depth
        t = rng.normal(0, 0.1, 3)
    x2 = K @ (R @ X + t[:, None])
    p2 = (x2[:2] / x2[2]).T + rng.normal(0, 0.35, (n, 2))
    out = rng.random(n) < rng.uniform(0.02, 0.25)                                       # gross outliers
    p2[out] += rng.normal(0, 25, (int(out.sum()), 2))
    status = (rng.random(n) < 0.9).astype(np.uint8)                                     # some features already dropped
    p1, p2 = p1.astype(np.float32), p2.astype(np.float32)
    m = status == 1
    cv2.setRNGSeed(seed)
    H, _ = cv2.findHomography(p1[m], p2[m], cv2.RANSAC, 3)
    F, _ = cv2.findFundamentalMat(p1[m], p2[m], cv2.FM_RANSAC, 3., 0.99)
    if H is None or F is None or F.shape != (3, 3):
        return None
    return dict(p1=p1, p2=p2, status=status, H=H.astype(np.float64), F=F.astype(np.float64), kind=kind)


def main():
    cases, seed = [], 5000
    while len(cases) < 60:
        c = make_case(seed)
        seed += 1
        if c is not None:
            cases.append(c)
    out = {"n": len(cases), "cv2_version": cv2.__version__}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "ransac_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()

"""Build recipe and ctypes loader of oracle/_ref/libpagk_ref.so.  TEST INFRASTRUCTURE ONLY.

libpagk_ref.so is the reference's OWN src/gyro_aided_tracker.cpp, src/patch_match.cpp and src/utils.cpp,
compiled unmodified from /root/reference (where they lie; nothing is copied into this repository) against the
stand-in OpenCV / Eigen3 / glog headers under oracle/ref_shim/ plus oracle/ref_harness.cpp.  What is real
and what is restated is spelled out at the top of oracle/ref_shim/pagk_cv_shim.hpp.

/root/reference exists only in the build container: `build()` compiles there; on the GPU box the prebuilt
library travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored) and `available()` says
whether it is present.  Imported by tests/ and bench.py's CPU legs only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PAGK_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(_HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libpagk_ref.so")
REF_SOURCES = ["src/gyro_aided_tracker.cpp", "src/patch_match.cpp", "src/utils.cpp", "src/frame.cpp", "src/ORBextractor.cc"]
# the reference's own flags (CMakeLists.txt:10-11, 17-20: -O3 -std=c++11, no -march, no -ffast-math)
CXXFLAGS = ["-O3", "-std=c++11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w"]
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def sources_present() -> bool:
    return all(os.path.exists(os.path.join(REFERENCE_ROOT, s)) for s in REF_SOURCES)


def available() -> bool:
    return os.path.exists(LIB)


def _deps():
    shim = os.path.join(_HERE, "ref_shim")
    own = [os.path.join(_HERE, "ref_harness.cpp"), os.path.join(_HERE, "pagk_cv_resize.h"), os.path.join(_HERE, "pagk_cv_fast.h"),
           os.path.join(shim, "pagk_cv_shim.hpp"), os.path.join(shim, "pagk_eigen_shim.hpp"),
           os.path.join(_HERE, "..", "include", "pagk.h")]
    return own + [os.path.join(REFERENCE_ROOT, s) for s in REF_SOURCES]


def build(force: bool = False) -> str | None:
    """compile the reference sources where they lie; no-op (returns the prebuilt path or None) without them"""
    if not sources_present():
        return LIB if available() else None
    if not force and available() and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in _deps()):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    shim = os.path.join(_HERE, "ref_shim")
    cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", shim, "-I", os.path.join(shim, "anchor"),
                                "-I", os.path.join(REFERENCE_ROOT, "include"), "-I", _HERE, "-o", LIB,
                                os.path.join(_HERE, "ref_harness.cpp")] + \
          [os.path.join(REFERENCE_ROOT, s) for s in REF_SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference build failed:\n" + r.stderr[-4000:])
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not available():
            build()
        if not available():
            raise RuntimeError("oracle/_ref/libpagk_ref.so is absent and /root/reference is not here to build it")
        lib = C.CDLL(LIB)
        lib.pagk_ref_track.argtypes = [C.POINTER(capi.PagkParams), C.POINTER(capi.PagkPairIn),
                                       C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_ref_track_batch.argtypes = [C.POINTER(capi.PagkParams), C.c_int, C.POINTER(capi.PagkPairIn),
                                             C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_ref_patch_match.argtypes = [C.POINTER(capi.PagkPatchMatchIn), C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_ref_integrate_gyro.argtypes = [C.POINTER(capi.PagkPairIn), _f32p, _f32p]
        lib.pagk_ref_inject_models.argtypes = [_f64p, _f64p]
        lib.pagk_ref_inject_models.restype = None
        lib.pagk_ref_geometry_validation.argtypes = [C.c_int, C.POINTER(capi.PagkGeometryIn), C.POINTER(capi.PagkGeometryOut)]
        lib.pagk_ref_set_predict_keypoints_and_mask.argtypes = [C.c_int, C.POINTER(capi.PagkCarryIn), C.POINTER(capi.PagkCarryOut)]
        lib.pagk_ref_orb_detect.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                            C.c_int, _f32p, _f32p, C.POINTER(C.c_int)]
        lib.pagk_ref_distribute_octtree.argtypes = [C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                    C.POINTER(C.c_int)]
        _lib = lib
    return _lib


def distribute_octtree(xy, response, min_x, max_x, min_y, max_y, n_features):
    """ORBextractor::DistributeOctTree (src/ORBextractor.cc:563-787) on candidates relative to (min_x, min_y): kept indices"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    rs = np.ascontiguousarray(response, np.float32).reshape(-1)
    idx, n = np.zeros(max(1, len(rs)), np.int32), C.c_int(0)
    load().pagk_ref_distribute_octtree(len(rs), xy.ctypes.data_as(_f32p), rs.ctypes.data_as(_f32p), int(min_x), int(max_x), int(min_y),
                                       int(max_y), int(n_features), idx.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n))
    return idx[:n.value].copy()


def last_path() -> int:
    """0: the call went through GyroAidedTracker::TrackFeatures() itself; 1: composed from the public pieces"""
    return int(load().pagk_ref_last_path())


def integrate_gyro(pair: capi.PairInputs):
    s = pair.as_struct()
    R = np.zeros(9, np.float32)
    M = np.zeros(9, np.float32)
    load().pagk_ref_integrate_gyro(C.byref(s), R.ctypes.data_as(_f32p), M.ctypes.data_as(_f32p))
    return R.reshape(3, 3), M.reshape(3, 3)


def patch_match(pm_struct, n_keys: int, n_threads: int = 1):
    out = capi.PairOutputs(n_keys)
    rc = load().pagk_ref_patch_match(C.byref(pm_struct), C.byref(out.struct), n_threads)
    return rc, out


def geometry_validation(cases):
    """GyroAidedTracker::GeometryValidation() with caller-supplied models; returns (rc, [PagkGeometryOut]), status in case.out_status"""
    ins = (capi.PagkGeometryIn * len(cases))()
    outs = (capi.PagkGeometryOut * len(cases))()
    for k, c in enumerate(cases):
        ins[k], outs[k] = c.structs()
    rc = load().pagk_ref_geometry_validation(len(cases), ins, outs)
    return rc, list(outs)


def set_predict_keypoints_and_mask(cases):
    """Frame::SetPredictKeyPointsAndMask(); returns (rc, [n_out]); the vectors land in each case's arrays"""
    ins = (capi.PagkCarryIn * len(cases))()
    outs = (capi.PagkCarryOut * len(cases))()
    for k, c in enumerate(cases):
        ins[k], outs[k] = c.structs()
    rc = load().pagk_ref_set_predict_keypoints_and_mask(len(cases), ins, outs)
    return rc, [int(o.n_out) for o in outs]


def orb_detect(img, ini_th=20, min_th=7, nfeatures=1000000, mask=None, max_out=400000):
    """ORBextractor(nfeatures, 1.2, 1, ini_th, min_th).DetectFeatures(img, mask): (xy [n][2], response [n]) in its output order"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    u8 = C.POINTER(C.c_uint8)
    xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    load().pagk_ref_orb_detect(img.ctypes.data_as(u8), w, h, img.strides[0], None if m is None else m.ctypes.data_as(u8), nfeatures,
                               ini_th, min_th, max_out, xy.ctypes.data_as(_f32p), rs.ctypes.data_as(_f32p), C.byref(n))
    k = min(n.value, max_out)
    return xy[:k].copy(), rs[:k].copy()


def track(pair: capi.PairInputs, params: capi.PagkParams, n_threads: int = 1):
    s = pair.as_struct()
    out = capi.PairOutputs(pair.n_keys)
    rc = load().pagk_ref_track(C.byref(params), C.byref(s), C.byref(out.struct), n_threads)
    return rc, out


def track_batch(pairs, params: capi.PagkParams, n_threads: int = 1, outs=None):
    ins = capi.make_in_array(pairs)
    outs = outs or [capi.PairOutputs(p.n_keys) for p in pairs]
    oarr = capi.make_out_array(outs)
    rc = load().pagk_ref_track_batch(C.byref(params), len(pairs), ins, oarr, n_threads)
    capi.sync_out_array(oarr, outs)
    return rc, outs


if __name__ == "__main__":
    print(build(force=True))

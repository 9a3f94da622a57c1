"""ctypes loader of the CPU oracle (oracle/libpagk_oracle.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) and by
nothing under the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libpagk_oracle.so")
_u8p, _f32p, _f64p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    # make knows the dependencies (the source, its headers and include/pagk.h: a changed struct must rebuild the checker)
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.pagk_oracle_resize_half.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.pagk_oracle_pyramid_level.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        lib.pagk_oracle_get_pixel_value.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        lib.pagk_oracle_get_pixel_value.restype = C.c_float
        lib.pagk_oracle_llt_solve.argtypes = [_f64p, _f64p, _f64p]
        lib.pagk_oracle_llt_solve.restype = None
        lib.pagk_oracle_set_llt_variant.argtypes = [C.c_int]
        lib.pagk_oracle_set_llt_variant.restype = None
        lib.pagk_oracle_affine_from_corners.argtypes = [_f32p, C.c_int, _f32p]
        lib.pagk_oracle_affine_from_corners.restype = None
        lib.pagk_oracle_integrate_gyro.argtypes = [C.POINTER(capi.PagkPairIn), _f32p, _f32p]
        lib.pagk_oracle_gyro_predict.argtypes = [C.POINTER(capi.PagkParams), C.POINTER(capi.PagkPairIn),
                                                 C.POINTER(capi.PagkPairOut)]
        lib.pagk_oracle_patch_match.argtypes = [C.POINTER(capi.PagkPatchMatchIn), C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_oracle_track.argtypes = [C.POINTER(capi.PagkParams), C.POINTER(capi.PagkPairIn),
                                          C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_oracle_track_batch.argtypes = [C.POINTER(capi.PagkParams), C.c_int, C.POINTER(capi.PagkPairIn),
                                                C.POINTER(capi.PagkPairOut), C.c_int]
        lib.pagk_oracle_geometry_validation.argtypes = [C.c_int, C.POINTER(capi.PagkGeometryIn), C.POINTER(capi.PagkGeometryOut)]
        lib.pagk_oracle_set_predict_keypoints_and_mask.argtypes = [C.c_int, C.POINTER(capi.PagkCarryIn), C.POINTER(capi.PagkCarryOut)]
        lib.pagk_oracle_fast_detect.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, _f32p,
                                                C.POINTER(C.c_int)]
        lib.pagk_oracle_orb_cell_detect.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, _f32p,
                                                    C.POINTER(C.c_int)]
        lib.pagk_oracle_remap_linear.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _u8p]
        _lib = lib
    return _lib


def level_size(width, height, level):
    c, r = width, height
    for _ in range(level):
        c, r = int(c * 0.5), int(r * 0.5)
    return c, r


def resize_half(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((int(h * 0.5), int(w * 0.5)), np.uint8)
    load().pagk_oracle_resize_half(img.ctypes.data_as(_u8p), w, h, img.strides[0], out.ctypes.data_as(_u8p))
    return out


def pyramid_level(img: np.ndarray, level: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    c, r = level_size(w, h, level)
    out = np.zeros((r, c), np.uint8)
    load().pagk_oracle_pyramid_level(img.ctypes.data_as(_u8p), w, h, img.strides[0], level, out.ctypes.data_as(_u8p))
    return out


def get_pixel_value(img: np.ndarray, x: float, y: float) -> float:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    return float(load().pagk_oracle_get_pixel_value(img.ctypes.data_as(_u8p), w, h, img.strides[0], x, y))


def set_llt_variant(v: int) -> None:
    """sensitivity study only (tools/eigen_variants.py): 0 = the oracle proper, 1..6 = one other plausible association"""
    load().pagk_oracle_set_llt_variant(int(v))


def llt_solve(H: np.ndarray, b: np.ndarray) -> np.ndarray:
    H = np.ascontiguousarray(H, np.float64).reshape(4, 4)
    b = np.ascontiguousarray(b, np.float64).reshape(4)
    x = np.zeros(4, np.float64)
    load().pagk_oracle_llt_solve(H.ctypes.data_as(_f64p), b.ctypes.data_as(_f64p), x.ctypes.data_as(_f64p))
    return x


def affine_from_corners(cflows: np.ndarray, half: int) -> np.ndarray:
    c = np.ascontiguousarray(cflows, np.float32).reshape(8)
    A = np.zeros(4, np.float32)
    load().pagk_oracle_affine_from_corners(c.ctypes.data_as(_f32p), half, A.ctypes.data_as(_f32p))
    return A.reshape(2, 2)


def integrate_gyro(pair: capi.PairInputs):
    s = pair.as_struct()
    R = np.zeros(9, np.float32)
    M = np.zeros(9, np.float32)
    load().pagk_oracle_integrate_gyro(C.byref(s), R.ctypes.data_as(_f32p), M.ctypes.data_as(_f32p))
    return R.reshape(3, 3), M.reshape(3, 3)


def gyro_predict(pair: capi.PairInputs, params: capi.PagkParams) -> capi.PairOutputs:
    s = pair.as_struct()
    out = capi.PairOutputs(pair.n_keys)
    rc = load().pagk_oracle_gyro_predict(C.byref(params), C.byref(s), C.byref(out.struct))
    assert rc == 0, rc
    return out


def patch_match(pm_struct, n_keys: int, n_threads: int = 1):
    out = capi.PairOutputs(n_keys)
    rc = load().pagk_oracle_patch_match(C.byref(pm_struct), C.byref(out.struct), n_threads)
    return rc, out


def geometry_validation(cases):
    """GyroAidedTracker::GeometryValidation() with caller-supplied models; returns (rc, [PagkGeometryOut]), status in case.out_status"""
    ins = (capi.PagkGeometryIn * len(cases))()
    outs = (capi.PagkGeometryOut * len(cases))()
    for k, c in enumerate(cases):
        ins[k], outs[k] = c.structs()
    rc = load().pagk_oracle_geometry_validation(len(cases), ins, outs)
    return rc, list(outs)


def set_predict_keypoints_and_mask(cases):
    """Frame::SetPredictKeyPointsAndMask(); returns (rc, [n_out]); the vectors land in each case's arrays"""
    ins = (capi.PagkCarryIn * len(cases))()
    outs = (capi.PagkCarryOut * len(cases))()
    for k, c in enumerate(cases):
        ins[k], outs[k] = c.structs()
    rc = load().pagk_oracle_set_predict_keypoints_and_mask(len(cases), ins, outs)
    return rc, [int(o.n_out) for o in outs]


def fast_detect(img, threshold, nonmax=True, mask=None, max_out=200000):
    """cv::FAST TYPE_9_16; returns (xy [n][2], response [n])"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    load().pagk_oracle_fast_detect(img.ctypes.data_as(_u8p), w, h, img.strides[0], threshold, 1 if nonmax else 0,
                                   None if m is None else m.ctypes.data_as(_u8p), max_out, xy.ctypes.data_as(_f32p),
                                   rs.ctypes.data_as(_f32p), C.byref(n))
    k = min(n.value, max_out)
    return xy[:k].copy(), rs[:k].copy()


def orb_cell_detect(img, ini_th=20, min_th=7, mask=None, max_out=400000):
    """per-cell FAST of ORBextractor::ComputeKeyPointsOctTree (one level) + the mask filter; (xy [n][2], response [n]), cell by cell"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    load().pagk_oracle_orb_cell_detect(img.ctypes.data_as(_u8p), w, h, img.strides[0], ini_th, min_th,
                                       None if m is None else m.ctypes.data_as(_u8p), max_out, xy.ctypes.data_as(_f32p),
                                       rs.ctypes.data_as(_f32p), C.byref(n))
    k = min(n.value, max_out)
    return xy[:k].copy(), rs[:k].copy()


def remap_linear(img, map_x, map_y):
    """cv::remap(img, map_x, map_y, INTER_LINEAR), constant 0 border"""
    img = np.ascontiguousarray(img, np.uint8)
    mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
    h, w = img.shape
    dh, dw = mx.shape
    out = np.zeros((dh, dw), np.uint8)
    load().pagk_oracle_remap_linear(img.ctypes.data_as(_u8p), w, h, img.strides[0], mx.ctypes.data_as(_f32p), my.ctypes.data_as(_f32p),
                                    dw, dh, out.ctypes.data_as(_u8p))
    return out


def track(pair: capi.PairInputs, params: capi.PagkParams, n_threads: int = 1):
    s = pair.as_struct()
    out = capi.PairOutputs(pair.n_keys)
    rc = load().pagk_oracle_track(C.byref(params), C.byref(s), C.byref(out.struct), n_threads)
    return rc, out


def track_batch(pairs, params: capi.PagkParams, n_threads: int = 1, outs=None):
    ins = capi.make_in_array(pairs)
    outs = outs or [capi.PairOutputs(p.n_keys) for p in pairs]
    oarr = capi.make_out_array(outs)
    rc = load().pagk_oracle_track_batch(C.byref(params), len(pairs), ins, oarr, n_threads)
    capi.sync_out_array(oarr, outs)
    return rc, outs

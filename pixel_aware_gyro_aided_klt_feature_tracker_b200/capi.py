"""ctypes mirror of include/pagk.h (the C-ABI of the CUDA hot path).

The structures below are field-for-field copies of the POD structs in ``include/pagk.h``; the
function table is every symbol that header declares.  ``load()`` opens the in-tree
``libpagk_cuda.so`` and fails loudly when it is missing -- there is no CPU fallback behind this
module (the CPU oracle under ``oracle/`` is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

PAGK_OK = 0
PAGK_ERR_INVALID = -1
PAGK_ERR_UNSUPPORTED = -2
PAGK_ERR_NO_DEVICE = -3
PAGK_ERR_CUDA = -4
PAGK_ERR_NOMEM = -5

# GyroAidedTracker::eType (reference include/gyro_aided_tracker.h:55-63)
OPENCV_OPTICAL_FLOW_PYR_LK = 0
GYRO_PREDICT = 1
GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED = 2
GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION = 3
GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION = 4
IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION = 5
GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR = 6
# GyroAidedTracker::ePredictMethod (:65-68)
PIXEL_AWARE_PREDICTION = 1
SINGLE_HOMOGRAPHY = 2

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)


class PagkConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("device", "max_width", "max_height", "max_keys", "max_pairs",
                                       "max_imu", "max_levels", "max_half_patch")]


class PagkParams(C.Structure):
    _fields_ = [("e_type", C.c_int), ("predict_method", C.c_int), ("half_patch", C.c_int),
                ("iterations", C.c_int), ("pyramids", C.c_int), ("inverse", C.c_int), ("calc_ncc", C.c_int),
                ("lambda_", C.c_float), ("alpha", C.c_float), ("max_distance", C.c_int)]


class PagkPairIn(C.Structure):
    _fields_ = [("img_ref", _u8p), ("img_cur", _u8p), ("width", C.c_int), ("height", C.c_int),
                ("pitch", C.c_int), ("n_keys", C.c_int), ("keys_ref_un", _f32p), ("keys_ref", _f32p),
                ("n_imu", C.c_int), ("imu_t", _f64p), ("imu_w", _f32p), ("t_ref", C.c_double),
                ("t_cur", C.c_double), ("bias_g", C.c_float * 3), ("K", C.c_float * 9),
                ("dist", C.c_float * 5), ("n_dist", C.c_int), ("Rbc", C.c_float * 9),
                ("normalize_table", _f32p), ("Rcl_override", _f32p)]


class PagkPairOut(C.Structure):
    _fields_ = [("pt_predict_un", _f32p), ("pt_predict", _f32p), ("status", _u8p),
                ("pt_gyro_predict_un", _f32p), ("pt_gyro_predict", _f32p), ("flows_predict_un", _f32p),
                ("affine", _f32p), ("corner_flows", _f32p), ("pt_corners_un", _f32p), ("pt_corners", _f32p),
                ("pm_pt_un", _f32p), ("pm_pt", _f32p), ("pm_status", _u8p), ("pixel_error", _f64p),
                ("distance", _f64p), ("ncc", _f32p), ("iters", _i32p), ("Rcl", C.c_float * 9),
                ("KRKinv", C.c_float * 9), ("n_predict", C.c_int), ("n_iterations", C.c_int64),
                ("t_gyro_predict", C.c_float), ("t_opt_flow", C.c_float), ("t_filter", C.c_float)]


class PagkPatchMatchIn(C.Structure):
    _fields_ = [("img_ref", _u8p), ("img_cur", _u8p), ("width", C.c_int), ("height", C.c_int),
                ("pitch", C.c_int), ("n_keys", C.c_int), ("keys_ref_un", _f32p), ("pt_predict_un", _f32p),
                ("status", _u8p), ("affine", _f32p), ("K", C.c_float * 9), ("dist", C.c_float * 5),
                ("n_dist", C.c_int), ("half_patch", C.c_int), ("iterations", C.c_int), ("pyramids", C.c_int),
                ("has_gyro_predict_initial", C.c_int), ("inverse", C.c_int), ("consider_illumination", C.c_int),
                ("consider_affine_deformation", C.c_int), ("regularization_penalty", C.c_int),
                ("calc_ncc", C.c_int), ("lambda_", C.c_float), ("alpha", C.c_float), ("max_distance", C.c_int)]


class PagkGeometryIn(C.Structure):
    _fields_ = [("n_keys", C.c_int), ("keys_ref_un", _f32p), ("pt_predict_un", _f32p), ("status", _u8p),
                ("H21", C.c_double * 9), ("F21", C.c_double * 9), ("sigma", C.c_float), ("estimate", C.c_int),
                ("seed", C.c_uint), ("reserved", C.c_int)]


class PagkGeometryOut(C.Structure):
    _fields_ = [("status", _u8p), ("score_H", C.c_float), ("score_F", C.c_float), ("used_H", C.c_int),
                ("n_candidates", C.c_int), ("n_inlier", C.c_int), ("H21", C.c_double * 9), ("F21", C.c_double * 9)]


class GeometryCase:
    """One GeometryValidation() call in numpy form: correspondences, status, the two models; owns the output status."""

    def __init__(self, keys_ref_un, pt_predict_un, status, H21=None, F21=None, sigma=1.0, estimate=False, seed=1):
        # H21 / F21 None with estimate=True: the device estimates both models (include/pagk.h)
        self.estimate, self.seed = bool(estimate), int(seed)
        H21 = np.eye(3) if H21 is None else H21
        F21 = np.zeros((3, 3)) if F21 is None else F21
        self.keys_ref_un = None if keys_ref_un is None else np.ascontiguousarray(keys_ref_un, np.float32).reshape(-1, 2)
        self.pt_predict_un = None if pt_predict_un is None else np.ascontiguousarray(pt_predict_un, np.float32).reshape(-1, 2)
        self.status = None if status is None else np.ascontiguousarray(status, np.uint8).reshape(-1)
        self.H21 = np.ascontiguousarray(H21, np.float64).reshape(3, 3)
        self.F21 = np.ascontiguousarray(F21, np.float64).reshape(3, 3)
        self.sigma = float(sigma)
        self.n_keys = 0 if self.status is None else self.status.size
        self.out_status = np.zeros(max(self.n_keys, 1), np.uint8)

    def structs(self, n_keys=None):
        i, o = PagkGeometryIn(), PagkGeometryOut()
        n = self.n_keys if n_keys is None else n_keys
        if self.out_status.size < n:
            self.out_status = np.zeros(n, np.uint8)
        i.n_keys = n
        i.keys_ref_un, i.pt_predict_un, i.status = _ptr(self.keys_ref_un, _f32p), _ptr(self.pt_predict_un, _f32p), _ptr(self.status, _u8p)
        i.H21[:] = self.H21.reshape(-1).tolist()
        i.F21[:] = self.F21.reshape(-1).tolist()
        i.sigma = self.sigma
        i.estimate, i.seed, i.reserved = int(self.estimate), self.seed, 0
        o.status = _ptr(self.out_status, _u8p)
        return i, o


class PagkCarryIn(C.Structure):
    _fields_ = [("n_keys", C.c_int), ("pt_predict", _f32p), ("pt_predict_un", _f32p), ("status", _u8p),
                ("keys_normal_last", _f32p), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("t_cur", C.c_double), ("t_last", C.c_double), ("width", C.c_int), ("height", C.c_int)]


class PagkCarryOut(C.Structure):
    _fields_ = [("n_out", C.c_int), ("keys", _f32p), ("keys_un", _f32p), ("keys_normal", _f32p), ("index_in_last", _i32p),
                ("flow_velocity_last", _f32p), ("mask", _u8p)]


class CarryCase:
    """One Frame::SetPredictKeyPointsAndMask() call in numpy form; owns the output arrays."""

    def __init__(self, pt_predict, pt_predict_un, status, keys_normal_last, K, t_cur, t_last, width, height, n_keys=None,
                 want_mask=True):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32).reshape(-1, 2)
        self.pt_predict, self.pt_predict_un = f(pt_predict), f(pt_predict_un)
        self.status = None if status is None else np.ascontiguousarray(status, np.uint8).reshape(-1)
        self.keys_normal_last = f(keys_normal_last)
        self.n_keys = int(n_keys if n_keys is not None else self.keys_normal_last.shape[0])
        K = np.asarray(K, np.float32).reshape(3, 3)
        self.fx, self.fy, self.cx, self.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
        self.t_cur, self.t_last, self.width, self.height = float(t_cur), float(t_last), int(width), int(height)
        n = max(self.n_keys, 1)
        self.keys, self.keys_un, self.keys_normal = (np.zeros((n, 2), np.float32) for _ in range(3))
        self.flow_velocity_last = np.zeros((n, 2), np.float32)
        self.index_in_last = np.full(n, -7, np.int32)
        self.mask = np.zeros((height, width), np.uint8) if want_mask else None

    def structs(self):
        i, o = PagkCarryIn(), PagkCarryOut()
        i.n_keys = self.n_keys
        i.pt_predict, i.pt_predict_un, i.status = _ptr(self.pt_predict, _f32p), _ptr(self.pt_predict_un, _f32p), _ptr(self.status, _u8p)
        i.keys_normal_last = _ptr(self.keys_normal_last, _f32p)
        i.fx, i.fy, i.cx, i.cy = self.fx, self.fy, self.cx, self.cy
        i.t_cur, i.t_last, i.width, i.height = self.t_cur, self.t_last, self.width, self.height
        o.keys, o.keys_un, o.keys_normal = _ptr(self.keys, _f32p), _ptr(self.keys_un, _f32p), _ptr(self.keys_normal, _f32p)
        o.index_in_last, o.flow_velocity_last = _ptr(self.index_in_last, _i32p), _ptr(self.flow_velocity_last, _f32p)
        o.mask = _ptr(self.mask, _u8p)
        return i, o


#: every symbol include/pagk.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "pagk_default_params": (None, [C.POINTER(PagkParams)]),
    "pagk_version": (C.c_int, []),
    "pagk_last_error": (C.c_char_p, []),
    "pagk_device_count": (C.c_int, []),
    "pagk_create": (C.c_int, [C.POINTER(PagkConfig), C.POINTER(_H)]),
    "pagk_destroy": (None, [_H]),
    "pagk_track_batch": (C.c_int, [_H, C.POINTER(PagkParams), C.c_int, C.POINTER(PagkPairIn), C.POINTER(PagkPairOut)]),
    "pagk_submit_batch": (C.c_int, [_H, C.POINTER(PagkParams), C.c_int, C.POINTER(PagkPairIn), C.POINTER(PagkPairOut)]),
    "pagk_wait_batch": (C.c_int, [_H]),
    "pagk_upload_batch": (C.c_int, [_H, C.POINTER(PagkParams), C.c_int, C.POINTER(PagkPairIn)]),
    "pagk_run_resident": (C.c_int, [_H]),
    "pagk_download_batch": (C.c_int, [_H, C.c_int, C.POINTER(PagkPairOut)]),
    "pagk_synchronize": (C.c_int, [_H]),
    "pagk_last_run_ms": (C.c_int, [_H] + [_f32p] * 5),
    "pagk_set_stage_timing": (C.c_int, [_H, C.c_int]),
    "pagk_set_device_share": (C.c_int, [_H, C.c_int]),
    "pagk_stream": (C.c_void_p, [_H]),
    "pagk_share_stream": (C.c_int, [_H, _H]),
    "pagk_timing_reset": (C.c_int, [_H]),
    "pagk_timing_read": (C.c_int, [_H, C.POINTER(C.c_int), _f32p]),
    "pagk_launch_count": (C.c_int64, [_H]),
    "pagk_build_pyramids": (C.c_int, [_H, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "pagk_pyramid_level_size": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pagk_get_pyramid_level": (C.c_int, [_H, C.c_int, C.c_int, _u8p, C.c_size_t]),
    "pagk_integrate_gyro": (C.c_int, [C.POINTER(PagkPairIn), _f32p, _f32p]),
    "pagk_gyro_predict": (C.c_int, [_H, C.POINTER(PagkParams), C.POINTER(PagkPairIn), C.POINTER(PagkPairOut)]),
    "pagk_patch_match": (C.c_int, [_H, C.POINTER(PagkPatchMatchIn), C.POINTER(PagkPairOut)]),
    "pagk_geometry_validation": (C.c_int, [_H, C.c_int, C.POINTER(PagkGeometryIn), C.POINTER(PagkGeometryOut)]),
    "pagk_fast_detect": (C.c_int, [_H, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, _f32p,
                                   C.POINTER(C.c_int)]),
    "pagk_orb_cell_detect": (C.c_int, [_H, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, _f32p,
                                       C.POINTER(C.c_int)]),
    "pagk_distribute_octtree": (C.c_int, [C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]),
    "pagk_orb_detect_features": (C.c_int, [_H, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, _f32p,
                                           C.POINTER(C.c_int)]),
    "pagk_set_rectify_maps": (C.c_int, [_H, _f32p, _f32p, C.c_int, C.c_int]),
    "pagk_remap_linear": (C.c_int, [_H, _u8p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _u8p]),
    "pagk_set_predict_keypoints_and_mask": (C.c_int, [_H, C.c_int, C.POINTER(PagkCarryIn), C.POINTER(PagkCarryOut)]),
}

LIB_NAME = "libpagk_cuda.so"


def lib_path() -> str:
    """the in-tree library; PAGK_LIB=<path> selects another build of the same ABI (kernel A/B runs)"""
    return os.environ.get("PAGK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)


_lib = None


def load(path: Optional[str] = None):
    """dlopen the CUDA library and type every symbol of include/pagk.h.  Raises when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or lib_path()
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the pagk hot path.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def default_params(**kw) -> PagkParams:
    """The reference's hard-coded values (src/gyro_aided_tracker.cpp:276-278, src/patch_match.cpp:48-50)."""
    p = PagkParams(e_type=GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                   predict_method=PIXEL_AWARE_PREDICTION, half_patch=5, iterations=10, pyramids=3,
                   inverse=0, calc_ncc=0, lambda_=1.0, alpha=0.5, max_distance=25)
    for k, v in kw.items():
        setattr(p, "lambda_" if k == "lambda" else k, v)
    return p


# ------------------------------------------------------------------------------------------------
# numpy <-> struct helpers shared by the product binding, the tests and the bench
# ------------------------------------------------------------------------------------------------
def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


class PairInputs:
    """One frame pair in numpy form (what the GyroAidedTracker ctor binds)."""

    def __init__(self, img_ref, img_cur, keys_ref_un, imu_t, imu_w, t_ref, t_cur, K, Rbc, dist=(0, 0, 0, 0, 0),
                 n_dist=4, bias_g=(0, 0, 0), keys_ref=None, normalize_table=None, Rcl_override=None):
        # img_ref None: stream continuation (the previous batch's current image of this pair index is the reference)
        self.img_ref = None if img_ref is None else np.ascontiguousarray(img_ref, dtype=np.uint8)
        self.img_cur = np.ascontiguousarray(img_cur, dtype=np.uint8)
        assert self.img_cur.ndim == 2 and (self.img_ref is None or self.img_ref.shape == self.img_cur.shape)
        self.keys_ref_un = np.ascontiguousarray(keys_ref_un, dtype=np.float32).reshape(-1, 2)
        self.keys_ref = (self.keys_ref_un if keys_ref is None
                         else np.ascontiguousarray(keys_ref, dtype=np.float32).reshape(-1, 2))
        self.imu_t = np.ascontiguousarray(imu_t, dtype=np.float64).reshape(-1)
        self.imu_w = np.ascontiguousarray(imu_w, dtype=np.float32).reshape(-1, 3)
        self.t_ref, self.t_cur = float(t_ref), float(t_cur)
        self.K = np.ascontiguousarray(K, dtype=np.float32).reshape(3, 3)
        self.Rbc = np.ascontiguousarray(Rbc, dtype=np.float32).reshape(3, 3)
        d = np.zeros(5, np.float32)
        dist = np.asarray(dist, np.float32).reshape(-1)
        d[:dist.size] = dist
        self.dist, self.n_dist = d, int(n_dist)
        self.bias_g = np.asarray(bias_g, np.float32).reshape(3)
        self.normalize_table = (None if normalize_table is None
                                else np.ascontiguousarray(normalize_table, dtype=np.float32))
        self.Rcl_override = (None if Rcl_override is None
                             else np.ascontiguousarray(Rcl_override, dtype=np.float32).reshape(3, 3))

    @property
    def n_keys(self):
        return self.keys_ref_un.shape[0]

    def as_struct(self) -> PagkPairIn:
        h, w = self.img_cur.shape
        s = PagkPairIn()
        s.img_ref = _ptr(self.img_ref, _u8p)
        s.img_cur = _ptr(self.img_cur, _u8p)
        s.width, s.height, s.pitch = w, h, self.img_cur.strides[0]
        s.n_keys = self.n_keys
        s.keys_ref_un = _ptr(self.keys_ref_un, _f32p)
        s.keys_ref = _ptr(self.keys_ref, _f32p)
        s.n_imu = self.imu_t.shape[0]
        s.imu_t = _ptr(self.imu_t, _f64p)
        s.imu_w = _ptr(self.imu_w, _f32p)
        s.t_ref, s.t_cur = self.t_ref, self.t_cur
        s.bias_g[:] = self.bias_g.tolist()
        s.K[:] = self.K.reshape(-1).tolist()
        s.dist[:] = self.dist.tolist()
        s.n_dist = self.n_dist
        s.Rbc[:] = self.Rbc.reshape(-1).tolist()
        s.normalize_table = _ptr(self.normalize_table, _f32p)
        s.Rcl_override = _ptr(self.Rcl_override, _f32p)
        return s


_OUT_SPEC = [  # (field, dtype, trailing shape)
    ("pt_predict_un", np.float32, (2,)), ("pt_predict", np.float32, (2,)), ("status", np.uint8, ()),
    ("pt_gyro_predict_un", np.float32, (2,)), ("pt_gyro_predict", np.float32, (2,)),
    ("flows_predict_un", np.float32, (2,)), ("affine", np.float32, (4,)), ("corner_flows", np.float32, (4, 2)),
    ("pt_corners_un", np.float32, (4, 2)), ("pt_corners", np.float32, (4, 2)), ("pm_pt_un", np.float32, (2,)),
    ("pm_pt", np.float32, (2,)), ("pm_status", np.uint8, ()), ("pixel_error", np.float64, ()),
    ("distance", np.float64, ()), ("ncc", np.float32, ()), ("iters", np.int32, ()),
]
_PTR_OF = {np.float32: _f32p, np.uint8: _u8p, np.float64: _f64p, np.int32: _i32p}


class PairOutputs:
    """Caller-allocated result vectors of one pair (the public members of GyroAidedTracker)."""

    def __init__(self, n_keys: int):
        self.n_keys = n_keys
        self.struct = PagkPairOut()
        for name, dt, tail in _OUT_SPEC:
            a = np.zeros((n_keys,) + tail, dtype=dt)
            setattr(self, name, a)
            setattr(self.struct, name, _ptr(a, _PTR_OF[dt]))

    @property
    def Rcl(self):
        return np.array(self.struct.Rcl[:], np.float32).reshape(3, 3)

    @property
    def KRKinv(self):
        return np.array(self.struct.KRKinv[:], np.float32).reshape(3, 3)

    @property
    def n_predict(self):
        return int(self.struct.n_predict)

    @property
    def n_iterations(self):
        return int(self.struct.n_iterations)

    def arrays(self):
        return {name: getattr(self, name) for name, _, _ in _OUT_SPEC}


def make_in_array(pairs: Sequence[PairInputs]):
    arr = (PagkPairIn * len(pairs))()
    for i, p in enumerate(pairs):
        arr[i] = p.as_struct()
    return arr


def make_out_array(outs: Sequence[PairOutputs]):
    arr = (PagkPairOut * len(outs))()
    for i, o in enumerate(outs):
        arr[i] = o.struct
    return arr


def sync_out_array(arr, outs: Sequence[PairOutputs]):
    """copy the scalar fields the callee wrote into the array elements back to the PairOutputs"""
    for i, o in enumerate(outs):
        o.struct = arr[i]


def patch_match_struct(p: PairInputs, pt_predict_un, status, affine, half_patch=5, iterations=10, pyramids=3,
                       has_gyro_predict_initial=1, inverse=0, consider_illumination=1,
                       consider_affine_deformation=1, regularization_penalty=0, calc_ncc=0, lambda_=1.0,
                       alpha=0.5, max_distance=25):
    """PatchMatch ctor arguments (include/patch_match.h:44-49) + what it reads from the tracker."""
    keep = dict(pt_predict_un=np.ascontiguousarray(pt_predict_un, np.float32),
                status=np.ascontiguousarray(status, np.uint8), affine=np.ascontiguousarray(affine, np.float32))
    h, w = p.img_ref.shape
    s = PagkPatchMatchIn()
    s.img_ref, s.img_cur = _ptr(p.img_ref, _u8p), _ptr(p.img_cur, _u8p)
    s.width, s.height, s.pitch, s.n_keys = w, h, p.img_ref.strides[0], p.n_keys
    s.keys_ref_un = _ptr(p.keys_ref_un, _f32p)
    s.pt_predict_un = _ptr(keep["pt_predict_un"], _f32p)
    s.status = _ptr(keep["status"], _u8p)
    s.affine = _ptr(keep["affine"], _f32p)
    s.K[:] = p.K.reshape(-1).tolist()
    s.dist[:] = p.dist.tolist()
    s.n_dist = p.n_dist
    s.half_patch, s.iterations, s.pyramids = half_patch, iterations, pyramids
    s.has_gyro_predict_initial, s.inverse = has_gyro_predict_initial, inverse
    s.consider_illumination, s.consider_affine_deformation = consider_illumination, consider_affine_deformation
    s.regularization_penalty, s.calc_ncc = regularization_penalty, calc_ncc
    s.lambda_, s.alpha, s.max_distance = lambda_, alpha, max_distance
    return s, keep

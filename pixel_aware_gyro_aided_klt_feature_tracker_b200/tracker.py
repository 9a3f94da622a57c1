"""Host-side mirror of the reference interface for the hot path, over the C-ABI (include/pagk.h).

``GyroAidedTracker`` and ``PatchMatch`` keep the reference's method names, constructor argument
order/meaning and public result members (reference include/gyro_aided_tracker.h:109-259,
include/patch_match.h:44-69) with numpy arrays where the reference has cv::Mat / std::vector.
All compute goes through ``libpagk_cuda.so``; without a CUDA device every call raises ``PagkError``
(there is no CPU path here -- the CPU oracle lives under ``oracle/`` and is test infrastructure).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import capi


class PagkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pagk error {code}: {msg}")
        self.code = code


def _check(lib, rc: int):
    if rc != capi.PAGK_OK:
        raise PagkError(rc, lib.pagk_last_error().decode("utf-8", "replace"))


def distribute_octtree(xy, response, min_x, max_x, min_y, max_y, n_features, lib=None):
    """ORBextractor::DistributeOctTree on candidates relative to (min_x, min_y): indices of the kept ones (host code, no device)"""
    lib = lib or capi.load()
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    rs = np.ascontiguousarray(response, np.float32).reshape(-1)
    idx, n = np.zeros(max(1, len(rs)), np.int32), C.c_int(0)
    f32 = C.POINTER(C.c_float)
    _check(lib, lib.pagk_distribute_octtree(len(rs), xy.ctypes.data_as(f32), rs.ctypes.data_as(f32), int(min_x), int(max_x), int(min_y),
                                            int(max_y), int(n_features), idx.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n)))
    return idx[:n.value].copy()


class Context:
    """One pagk_handle = the device workspace of one GPU (images, keypoints, results)."""

    def __init__(self, device: int = 0, max_width: int = 752, max_height: int = 480, max_keys: int = 1024,
                 max_pairs: int = 64, max_imu: int = 64, max_levels: int = 4, max_half_patch: int = 5):
        self.lib = capi.load()
        self.cfg = capi.PagkConfig(device=device, max_width=max_width, max_height=max_height, max_keys=max_keys,
                                   max_pairs=max_pairs, max_imu=max_imu, max_levels=max_levels,
                                   max_half_patch=max_half_patch)
        self.handle = C.c_void_p()
        _check(self.lib, self.lib.pagk_create(C.byref(self.cfg), C.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.pagk_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- whole path -------------------------------------------------------------------------
    def track_batch(self, pairs: Sequence[capi.PairInputs], params: capi.PagkParams,
                    outs: Optional[List[capi.PairOutputs]] = None) -> List[capi.PairOutputs]:
        ins = capi.make_in_array(pairs)
        outs = outs or [capi.PairOutputs(p.n_keys) for p in pairs]
        oarr = capi.make_out_array(outs)
        rc = self.lib.pagk_track_batch(self.handle, C.byref(params), len(pairs), ins, oarr)
        capi.sync_out_array(oarr, outs)
        _check(self.lib, rc)
        return outs

    def submit_prepared(self, params, ins, oarr, n):
        """asynchronous pagk_track_batch on prebuilt ctypes arrays (kept alive by the caller until wait())"""
        _check(self.lib, self.lib.pagk_submit_batch(self.handle, C.byref(params), n, ins, oarr))

    def wait(self):
        _check(self.lib, self.lib.pagk_wait_batch(self.handle))

    def upload(self, pairs, params):
        self._ins = capi.make_in_array(pairs)  # keep alive
        self._pairs = pairs
        _check(self.lib, self.lib.pagk_upload_batch(self.handle, C.byref(params), len(pairs), self._ins))

    def run(self):
        _check(self.lib, self.lib.pagk_run_resident(self.handle))

    def synchronize(self):
        _check(self.lib, self.lib.pagk_synchronize(self.handle))

    def download(self, outs: List[capi.PairOutputs]):
        oarr = capi.make_out_array(outs)
        rc = self.lib.pagk_download_batch(self.handle, len(outs), oarr)
        capi.sync_out_array(oarr, outs)
        _check(self.lib, rc)
        return outs

    def last_run_ms(self):
        v = [C.c_float() for _ in range(5)]
        _check(self.lib, self.lib.pagk_last_run_ms(self.handle, *[C.byref(x) for x in v]))
        return dict(zip(("total", "pyramid", "predict", "lk", "filter"), [x.value for x in v]))

    def geometry_validation(self, cases):
        """GyroAidedTracker::GeometryValidation() without its RANSAC estimators (include/pagk.h): a list of capi.GeometryCase
        (vectors None = the resident results of the last run); returns the PagkGeometryOut list, status in case.out_status"""
        ins = (capi.PagkGeometryIn * len(cases))()
        outs = (capi.PagkGeometryOut * len(cases))()
        for k, c in enumerate(cases):
            ins[k], outs[k] = c.structs(n_keys=c.n_keys if c.status is not None else getattr(c, "resident_n_keys", 0))
        _check(self.lib, self.lib.pagk_geometry_validation(self.handle, len(cases), ins, outs))
        return list(outs)

    def set_predict_keypoints_and_mask(self, cases):
        """Frame::SetPredictKeyPointsAndMask() (include/pagk.h): a list of capi.CarryCase (prediction vectors None = the
        resident results of the last run); returns the survivor counts, vectors and mask land in each case's arrays"""
        ins = (capi.PagkCarryIn * len(cases))()
        outs = (capi.PagkCarryOut * len(cases))()
        for k, c in enumerate(cases):
            ins[k], outs[k] = c.structs()
        _check(self.lib, self.lib.pagk_set_predict_keypoints_and_mask(self.handle, len(cases), ins, outs))
        return [int(o.n_out) for o in outs]

    def fast_detect(self, img, threshold, nonmax=True, mask=None, max_out=200000):
        """cv::FAST TYPE_9_16 (include/pagk.h); returns (xy [n][2], response [n]) in OpenCV's order"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        u8 = C.POINTER(C.c_uint8)
        f32 = C.POINTER(C.c_float)
        _check(self.lib, self.lib.pagk_fast_detect(self.handle, img.ctypes.data_as(u8), w, h, img.strides[0], int(threshold),
                                                   1 if nonmax else 0, None if m is None else m.ctypes.data_as(u8), max_out,
                                                   xy.ctypes.data_as(f32), rs.ctypes.data_as(f32), C.byref(n)))
        k = min(n.value, max_out)
        return xy[:k].copy(), rs[:k].copy()

    def orb_cell_detect(self, img, ini_th=20, min_th=7, mask=None, max_out=400000):
        """the per-cell FAST detection of ORBextractor::DetectFeatures, one level (include/pagk.h)"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        u8 = C.POINTER(C.c_uint8)
        f32 = C.POINTER(C.c_float)
        _check(self.lib, self.lib.pagk_orb_cell_detect(self.handle, img.ctypes.data_as(u8), w, h, img.strides[0], int(ini_th), int(min_th),
                                                       None if m is None else m.ctypes.data_as(u8), max_out, xy.ctypes.data_as(f32),
                                                       rs.ctypes.data_as(f32), C.byref(n)))
        k = min(n.value, max_out)
        return xy[:k].copy(), rs[:k].copy()

    def orb_detect_features(self, img, n_features, ini_th=20, min_th=7, mask=None, max_out=100000):
        """ORBextractor(n_features, 1.2, 1, ini_th, min_th).DetectFeatures(img, mask), one level (include/pagk.h)"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        xy, rs, n = np.zeros((max_out, 2), np.float32), np.zeros(max_out, np.float32), C.c_int(0)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        u8 = C.POINTER(C.c_uint8)
        f32 = C.POINTER(C.c_float)
        _check(self.lib, self.lib.pagk_orb_detect_features(self.handle, img.ctypes.data_as(u8), w, h, img.strides[0], int(n_features),
                                                           int(ini_th), int(min_th), None if m is None else m.ctypes.data_as(u8), max_out,
                                                           xy.ctypes.data_as(f32), rs.ctypes.data_as(f32), C.byref(n)))
        k = min(n.value, max_out)
        return xy[:k].copy(), rs[:k].copy()

    def remap_linear(self, img, map_x, map_y):
        """cv::remap(img, map_x, map_y, INTER_LINEAR) with the constant 0 border (include/pagk.h)"""
        img = np.ascontiguousarray(img, np.uint8)
        mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
        h, w = img.shape
        dh, dw = mx.shape
        out = np.zeros((dh, dw), np.uint8)
        u8, f32 = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        _check(self.lib, self.lib.pagk_remap_linear(self.handle, img.ctypes.data_as(u8), w, h, img.strides[0], mx.ctypes.data_as(f32),
                                                    my.ctypes.data_as(f32), dw, dh, out.ctypes.data_as(u8)))
        return out

    def set_rectify_maps(self, map_x, map_y):
        """batches then bring DISTORTED images, remapped on the device into their pyramid slots; (None, None) switches it off"""
        if map_x is None:
            _check(self.lib, self.lib.pagk_set_rectify_maps(self.handle, None, None, 0, 0))
            return
        mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
        f32 = C.POINTER(C.c_float)
        _check(self.lib, self.lib.pagk_set_rectify_maps(self.handle, mx.ctypes.data_as(f32), my.ctypes.data_as(f32), mx.shape[1], mx.shape[0]))

    def set_stage_timing(self, on: bool):
        """CUDA events between the kernels of a run (per-stage clocks); off for throughput pipelines"""
        _check(self.lib, self.lib.pagk_set_stage_timing(self.handle, 1 if on else 0))

    def set_device_share(self, n_handles: int):
        """this handle is one of n_handles that keep launches in flight on the device: a launch of the patch-alignment kernel
        takes 1/n_handles of every SM's CTA slots and the other handles' launches run beside it (pagk_set_device_share)"""
        _check(self.lib, self.lib.pagk_set_device_share(self.handle, int(n_handles)))

    def share_stream(self, other: "Context"):
        _check(self.lib, self.lib.pagk_share_stream(self.handle, other.handle))

    def timing_reset(self):
        _check(self.lib, self.lib.pagk_timing_reset(self.handle))

    def timing_read(self):
        n, ms = C.c_int(), C.c_float()
        _check(self.lib, self.lib.pagk_timing_read(self.handle, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self) -> int:
        return int(self.lib.pagk_launch_count(self.handle))

    # ---- stages -----------------------------------------------------------------------------
    def build_pyramids(self, images: Sequence[np.ndarray], levels: int):
        imgs = [np.ascontiguousarray(i, np.uint8) for i in images]
        h, w = imgs[0].shape
        arr = (C.POINTER(C.c_uint8) * len(imgs))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in imgs])
        _check(self.lib, self.lib.pagk_build_pyramids(self.handle, len(imgs), arr, w, h, imgs[0].strides[0], levels))
        self._pyr_wh = (w, h)

    def pyramid_level(self, image: int, level: int) -> np.ndarray:
        w, h = self._pyr_wh
        c, r = C.c_int(), C.c_int()
        _check(self.lib, self.lib.pagk_pyramid_level_size(w, h, level, C.byref(c), C.byref(r)))
        out = np.zeros((r.value, c.value), np.uint8)
        _check(self.lib, self.lib.pagk_get_pyramid_level(self.handle, image, level,
                                                         out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size))
        return out

    def gyro_predict(self, pair: capi.PairInputs, params: capi.PagkParams) -> capi.PairOutputs:
        s = pair.as_struct()
        out = capi.PairOutputs(pair.n_keys)
        _check(self.lib, self.lib.pagk_gyro_predict(self.handle, C.byref(params), C.byref(s), C.byref(out.struct)))
        return out

    def patch_match(self, pm_struct, n_keys: int) -> capi.PairOutputs:
        out = capi.PairOutputs(n_keys)
        _check(self.lib, self.lib.pagk_patch_match(self.handle, C.byref(pm_struct), C.byref(out.struct)))
        return out


def integrate_gyro(pair: capi.PairInputs):
    """IntegrateGyroMeasurements + SetRcl (host arithmetic of the product library)."""
    lib = capi.load()
    s = pair.as_struct()
    R = np.zeros(9, np.float32)
    M = np.zeros(9, np.float32)
    fp = C.POINTER(C.c_float)
    _check(lib, lib.pagk_integrate_gyro(C.byref(s), R.ctypes.data_as(fp), M.ctypes.data_as(fp)))
    return R.reshape(3, 3), M.reshape(3, 3)


# ------------------------------------------------------------------------------------------------
# Reference-shaped classes
# ------------------------------------------------------------------------------------------------
@dataclass
class CameraParams:
    """reference include/imu_types.h:34-88 (mK, mDistCoef 4x1, width, height)"""
    mK: np.ndarray
    mDistCoef: np.ndarray
    width: int
    height: int


@dataclass
class Frame:
    """the Frame fields the tracker reads (reference include/frame.h:59-91)"""
    mTimeStamp: float
    mGray: np.ndarray
    mvKeys: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    mvKeysUn: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    mvImuFromLastFrame: Optional[tuple] = None   # (t[n] float64, w[n][3] float32)
    mpCameraParams: Optional[CameraParams] = None
    # filled by SetBackToFrame
    mvPtGyroPredictUn: Optional[np.ndarray] = None
    mvPtPredict: Optional[np.ndarray] = None
    mvPtPredictUn: Optional[np.ndarray] = None
    mvStatus: Optional[np.ndarray] = None
    mvNcc: Optional[np.ndarray] = None
    mvvFlowsPredictCorners: Optional[np.ndarray] = None
    mRcl: Optional[np.ndarray] = None


class GyroAidedTracker:
    """GyroAidedTracker(pFrameRef, pFrameCur, imuCalib, biasg, normalizeTable, type, predictMethod,
    saveFolderPath, halfPatchSize) -- reference src/gyro_aided_tracker.cpp:30-49."""

    OPENCV_OPTICAL_FLOW_PYR_LK = capi.OPENCV_OPTICAL_FLOW_PYR_LK
    GYRO_PREDICT = capi.GYRO_PREDICT
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED = capi.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION = capi.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION = capi.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR = capi.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR
    IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION = capi.IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION
    PIXEL_AWARE_PREDICTION = capi.PIXEL_AWARE_PREDICTION
    SINGLE_HOMOGRAPHY = capi.SINGLE_HOMOGRAPHY

    def __init__(self, ctx: Context, pFrameRef: Frame, pFrameCur: Frame, Tbc: np.ndarray, biasg=(0, 0, 0),
                 normalizeTable: Optional[np.ndarray] = None,
                 type_: int = capi.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                 predictMethod_: int = capi.PIXEL_AWARE_PREDICTION, saveFolderPath: str = "", halfPatchSize_: int = 5,
                 iterations: int = 10, pyramids: int = 3):
        self._ctx = ctx
        self.mType, self.mPredictMethod = type_, predictMethod_
        self.mSaveFolderPath = saveFolderPath   # the reference only mkdir -p's it; nothing is written here
        self.mHalfPatchSize = 5 if halfPatchSize_ == 0 else halfPatchSize_
        cam = pFrameCur.mpCameraParams
        imu_t, imu_w = pFrameCur.mvImuFromLastFrame if pFrameCur.mvImuFromLastFrame is not None else (np.zeros(0), np.zeros((0, 3)))
        dist = np.asarray(cam.mDistCoef, np.float32).reshape(-1)
        self._pair = capi.PairInputs(pFrameRef.mGray, pFrameCur.mGray, pFrameRef.mvKeysUn, imu_t, imu_w,
                                     pFrameRef.mTimeStamp, pFrameCur.mTimeStamp, cam.mK,
                                     np.asarray(Tbc, np.float32)[:3, :3], dist=dist, n_dist=dist.size, bias_g=biasg,
                                     keys_ref=pFrameRef.mvKeys, normalize_table=normalizeTable)
        self.mN = self._pair.n_keys
        self._iterations, self._pyramids = iterations, pyramids
        self._regular_override = None
        self.mRcl = None
        self._out = capi.PairOutputs(self.mN)
        self._export()

    # -- reference method names ---------------------------------------------------------------
    def SetType(self, type_):
        self.mType = type_

    def SetRegularizationPenalty(self, flag: bool):
        # TrackFeatures() overwrites the flag from eType (reference :384-414), so this is a no-op there too
        self._regular_override = bool(flag)

    def SetRcl(self, Rcl):
        self._pair.Rcl_override = np.ascontiguousarray(Rcl, np.float32).reshape(3, 3)
        self.mRcl = self._pair.Rcl_override.copy()

    def GetRcl(self):
        return None if self.mRcl is None else self.mRcl.copy()

    def _params(self):
        return capi.default_params(e_type=self.mType, predict_method=self.mPredictMethod,
                                   half_patch=self.mHalfPatchSize, iterations=self._iterations,
                                   pyramids=self._pyramids)

    def TrackFeatures(self) -> int:
        try:
            self._ctx.track_batch([self._pair], self._params(), [self._out])
        except PagkError as e:
            if e.code == capi.PAGK_ERR_UNSUPPORTED:
                return -1   # "Unsupport type!!! return -1;" (reference :415-418)
            raise
        self._export()
        return self._out.n_predict

    def SetBackToFrame(self, pFrame: Frame):
        """reference src/gyro_aided_tracker.cpp:97-111"""
        pFrame.mvPtGyroPredictUn = self.mvPtGyroPredictUn.copy()
        pFrame.mvPtPredict = self.mvPtPredict.copy()
        pFrame.mvPtPredictUn = self.mvPtPredictUn.copy()
        pFrame.mvStatus = self.mvStatus.copy()
        pFrame.mvNcc = self.mvNccAfterPatchMatched.copy()
        pFrame.mvvFlowsPredictCorners = self.mvvFlowsPredictCorners.copy()
        pFrame.mRcl = None if self.mRcl is None else self.mRcl.copy()

    def _export(self):
        o = self._out
        self.mvPtPredict, self.mvPtPredictUn = o.pt_predict, o.pt_predict_un
        self.mvPtGyroPredict, self.mvPtGyroPredictUn = o.pt_gyro_predict, o.pt_gyro_predict_un
        self.mvvPtPredictCorners, self.mvvPtPredictCornersUn = o.pt_corners, o.pt_corners_un
        self.mvvFlowsPredictCorners = o.corner_flows
        self.mvStatus = o.status
        self.mvAffineDeformationMatrix = o.affine.reshape(-1, 2, 2)
        self.mvPtPredictAfterPatchMatched, self.mvPtPredictAfterPatchMatchedUn = o.pm_pt, o.pm_pt_un
        self.mvStatusAfterPatchMatched = o.pm_status
        self.mvPixelErrorsOfPatchMatched = o.pixel_error
        self.mvDistanceBetweenPredictedAndPatchMatched = o.distance
        self.mvNccAfterPatchMatched = o.ncc
        self.mvFlowsPredictUn = o.flows_predict_un
        self.mTimeCostGyroPredict = o.struct.t_gyro_predict
        self.mTimeCostOptFlow = o.struct.t_opt_flow
        self.mTimeCostOptFlowResultFilterOut = o.struct.t_filter
        if o.struct.n_predict or np.any(o.Rcl):
            self.mRcl = o.Rcl
            self.mKRKinv = o.KRKinv


class PatchMatch:
    """PatchMatch(pMatcher, halfPatchSize, iterations, pyramids, bHasGyroPredictInitial, bInverse,
    bConsiderIllumination, bConsiderAffineDeformation, bRegularizationPenalty=True, bCalculateNCC=False)
    -- reference include/patch_match.h:44-49.  OpticalFlowMultiLevel() writes the six SetMatcher
    vectors back into the tracker (reference src/patch_match.cpp:370-388)."""

    def __init__(self, pMatcher_: GyroAidedTracker, halfPatchSize_: int, iterations_: int, pyramids_: int,
                 bHasGyroPredictInitial_: bool, bInverse_: bool, bConsiderIllumination_: bool,
                 bConsiderAffineDeformation_: bool, bRegularizationPenalty_: bool = True, bCalculateNCC_: bool = False):
        self.mpMatcher = pMatcher_
        self._args = dict(half_patch=halfPatchSize_, iterations=iterations_, pyramids=pyramids_,
                          has_gyro_predict_initial=int(bHasGyroPredictInitial_), inverse=int(bInverse_),
                          consider_illumination=int(bConsiderIllumination_),
                          consider_affine_deformation=int(bConsiderAffineDeformation_),
                          regularization_penalty=int(bRegularizationPenalty_), calc_ncc=int(bCalculateNCC_))
        self.mvGyroPredictStatus = np.array(pMatcher_.mvStatus, np.uint8, copy=True)

    def OpticalFlowMultiLevel(self):
        t = self.mpMatcher
        s, keep = capi.patch_match_struct(t._pair, t.mvPtPredictUn, self.mvGyroPredictStatus,
                                          t.mvAffineDeformationMatrix.reshape(-1, 4), **self._args)
        out = t._ctx.patch_match(s, t.mN)
        del keep
        t.mvPtPredictAfterPatchMatched, t.mvPtPredictAfterPatchMatchedUn = out.pm_pt, out.pm_pt_un
        t.mvStatusAfterPatchMatched = out.pm_status
        t.mvPixelErrorsOfPatchMatched = out.pixel_error
        t.mvDistanceBetweenPredictedAndPatchMatched = out.distance
        t.mvNccAfterPatchMatched = out.ncc
        self.iters = out.iters

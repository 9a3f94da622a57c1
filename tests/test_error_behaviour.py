"""Error behaviour at the boundary (include/pagk.h): what the reference answers with -1 or an assert,
the C-ABI answers with a status code and a message; no exception crosses it, nothing is computed on the CPU instead, and
the handle stays usable.  Reference: TrackFeatures() returns -1 for an unsupported eType
(src/gyro_aided_tracker.cpp:415-418)."""
import copy
import ctypes as C

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_ctx(cuda_lib):
    if cuda_lib.pagk_device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible")
    with tracker.Context(max_width=320, max_height=240, max_keys=128, max_pairs=2, max_levels=3, max_half_patch=5,
                         max_imu=32) as ctx:
        yield ctx


@pytest.fixture(scope="module")
def pair():
    return synth.make_pair(7700, width=320, height=240, n_keys=100, pyramids=3, border=24)


def _code(ctx, pairs, prm):
    with pytest.raises(tracker.PagkError) as e:
        ctx.track_batch(pairs, prm)
    assert ctx.lib.pagk_last_error().decode() != ""
    return e.value.code


def test_unsupported_modes_answer_like_track_features(small_ctx, pair):
    for kw in (dict(e_type=capi.OPENCV_OPTICAL_FLOW_PYR_LK), dict(e_type=7), dict(e_type=-1), dict(inverse=1)):
        prm = capi.default_params(pyramids=3, **kw)
        outs = [capi.PairOutputs(pair.n_keys)]
        with pytest.raises(tracker.PagkError) as e:
            small_ctx.track_batch([pair], prm, outs)
        assert e.value.code == capi.PAGK_ERR_UNSUPPORTED, kw
        assert outs[0].n_predict == -1, kw                      # the reference's return value


def test_sizes_beyond_the_handle_limits(small_ctx, pair):
    prm = capi.default_params(pyramids=3)
    big = synth.make_pair(7701, width=352, height=240, n_keys=100, pyramids=3, border=24)
    assert _code(small_ctx, [big], prm) == capi.PAGK_ERR_INVALID                               # image larger than max_width
    many = synth.make_pair(7702, width=320, height=240, n_keys=129, pyramids=3, border=24)
    assert _code(small_ctx, [many], prm) == capi.PAGK_ERR_INVALID                              # n_keys > max_keys
    assert _code(small_ctx, [pair, pair, pair], prm) == capi.PAGK_ERR_INVALID                  # n_pairs > max_pairs
    assert _code(small_ctx, [pair], capi.default_params(pyramids=4)) == capi.PAGK_ERR_INVALID  # levels > max_levels
    assert _code(small_ctx, [pair], capi.default_params(pyramids=3, half_patch=6)) == capi.PAGK_ERR_INVALID
    assert _code(small_ctx, [pair], capi.default_params(pyramids=3, iterations=-1)) == capi.PAGK_ERR_INVALID
    assert _code(small_ctx, [pair], capi.default_params(pyramids=3, predict_method=9)) == capi.PAGK_ERR_INVALID
    other = synth.make_pair(7703, width=256, height=240, n_keys=100, pyramids=3, border=24)
    assert _code(small_ctx, [pair, other], prm) == capi.PAGK_ERR_INVALID                       # one geometry per batch


def test_null_pointers(small_ctx, pair):
    prm = capi.default_params(pyramids=3)
    lib, h = small_ctx.lib, small_ctx.handle
    ins, out = capi.make_in_array([pair]), capi.PairOutputs(pair.n_keys)
    oarr = capi.make_out_array([out])
    assert lib.pagk_track_batch(None, C.byref(prm), 1, ins, oarr) == capi.PAGK_ERR_INVALID
    assert lib.pagk_track_batch(h, None, 1, ins, oarr) == capi.PAGK_ERR_INVALID
    assert lib.pagk_track_batch(h, C.byref(prm), 1, None, oarr) == capi.PAGK_ERR_INVALID
    for field, typ in (("img_cur", capi._u8p), ("keys_ref_un", capi._f32p), ("imu_t", C.POINTER(C.c_double))):
        bad = capi.make_in_array([pair])
        setattr(bad[0], field, C.cast(None, typ))
        assert lib.pagk_track_batch(h, C.byref(prm), 1, bad, oarr) == capi.PAGK_ERR_INVALID, field
    bad = capi.make_in_array([pair])
    bad[0].pitch = pair.img_cur.shape[1] - 1
    assert lib.pagk_track_batch(h, C.byref(prm), 1, bad, oarr) == capi.PAGK_ERR_INVALID
    # eType 5 starts from the reference keypoints themselves: keys_ref must be there
    bad = capi.make_in_array([pair])
    bad[0].keys_ref = C.cast(None, capi._f32p)
    p5 = capi.default_params(pyramids=3, e_type=capi.IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION)
    assert lib.pagk_track_batch(h, C.byref(p5), 1, bad, oarr) == capi.PAGK_ERR_INVALID


def test_calls_out_of_order(cuda_lib):
    with tracker.Context(max_width=320, max_height=240, max_keys=128, max_pairs=2, max_levels=3) as ctx:
        assert ctx.lib.pagk_run_resident(ctx.handle) == capi.PAGK_ERR_INVALID          # nothing uploaded
        out = capi.PairOutputs(4)
        assert ctx.lib.pagk_download_batch(ctx.handle, 1, capi.make_out_array([out])) == capi.PAGK_ERR_INVALID
    cfg = capi.PagkConfig(device=99, max_width=320, max_height=240, max_keys=16, max_pairs=1, max_imu=16, max_levels=3, max_half_patch=5)
    h = C.c_void_p()
    assert cuda_lib.pagk_create(C.byref(cfg), C.byref(h)) == capi.PAGK_ERR_INVALID and not h
    cfg = capi.PagkConfig(device=0, max_width=0, max_height=240, max_keys=16, max_pairs=1, max_imu=16, max_levels=3, max_half_patch=5)
    assert cuda_lib.pagk_create(C.byref(cfg), C.byref(h)) == capi.PAGK_ERR_INVALID and not h
    assert cuda_lib.pagk_create(None, C.byref(h)) == capi.PAGK_ERR_INVALID


def test_the_handle_survives_its_errors(small_ctx, oracle, pair):
    prm = capi.default_params(pyramids=3)
    _code(small_ctx, [pair, pair, pair], prm)
    _code(small_ctx, [pair], capi.default_params(pyramids=3, e_type=0))
    (g,) = small_ctx.track_batch([pair], prm)
    rc, c = oracle.track(pair, prm, 1)
    assert rc == 0
    helpers.assert_bit_exact(g, c)
    empty = copy.copy(pair)
    empty.keys_ref_un = np.zeros((0, 2), np.float32)
    empty.keys_ref = empty.keys_ref_un
    (g0,) = small_ctx.track_batch([empty], prm)                  # no keypoints is not an error (the reference returns 0)
    assert g0.n_predict == 0


def test_non_finite_inputs_fail_per_feature(small_ctx, oracle, pair):
    """Repeated IMU timestamps (tab = 0, src/gyro_aided_tracker.cpp:531-535) or a NaN gyro sample make Rcl NaN; the reference
    (and the restatement, faithfully) then index images with int(NaN) and die with SIGSEGV.  The CUDA path answers with
    status 0 for every feature it cannot place, touches no memory it does not own, and the handle stays usable."""
    prm = capi.default_params(pyramids=3)
    for case in ("dup01", "dup12", "nan_w"):
        q = copy.copy(pair)
        if case == "nan_w":
            q.imu_w = pair.imu_w.copy(); q.imu_w[2, 1] = np.nan
        else:
            q.imu_t, q.imu_w = pair.imu_t[:3].copy(), pair.imu_w[:3].copy()
            k = 1 if case == "dup01" else 2
            q.imu_t[k] = q.imu_t[k - 1]
        (g,) = small_ctx.track_batch([q], prm)
        assert np.isnan(np.asarray(g.Rcl)).any(), case
        assert g.n_predict == 0 and not g.status.any(), case
    q = copy.copy(pair)
    q.keys_ref_un = pair.keys_ref_un.copy()
    q.keys_ref_un[3] = np.nan
    q.keys_ref_un[5, 0] = np.inf
    q.keys_ref = q.keys_ref_un
    (g,) = small_ctx.track_batch([q], prm)
    assert g.status[3] == 0 and g.status[5] == 0
    rc, c = oracle.track(pair, prm, 1)                            # the finite features are tracked as if the others were not there
    keep = np.ones(pair.n_keys, bool); keep[[3, 5]] = False
    assert np.array_equal(g.pm_status[keep], c.pm_status[keep]) and np.array_equal(g.pm_pt_un[keep].view(np.uint32), c.pm_pt_un[keep].view(np.uint32))
    (g,) = small_ctx.track_batch([pair], prm)
    helpers.assert_bit_exact(g, c)

"""The drivers' frame loop (Examples/Demo/RealSenseD435i.cpp:143-324) over a synthetic stream, piece by piece through each
backend: TrackFeatures -> SetPredictKeyPointsAndMask -> keypoint top-up (per-cell FAST under the occupancy mask) -> next
frame.  GeometryValidation's RANSAC and the octree thinning are left out (not reproducible); the top-up takes
the strongest candidates.  Every frame's outputs must agree bit for bit between the backends."""
import copy

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth
from tests import helpers

W, H, N_FRAMES, N_WANT = 320, 240, 6, 150


class OracleBackend:
    def __init__(self, mod):
        self.m = mod

    def track(self, pair, prm, first):
        rc, o = self.m.track(pair, prm, 2)
        assert rc == 0
        return o

    def carry(self, out, last_normal, pair):
        c = capi.CarryCase(out.pt_predict, out.pt_predict_un, out.status, last_normal, pair.K, pair.t_cur, pair.t_ref, W, H)
        rc, (n,) = self.m.set_predict_keypoints_and_mask([c])
        assert rc == 0
        return c, n

    def detect(self, img, mask):
        return self.m.orb_cell_detect(img, 20, 7, mask=mask)


class GpuBackend:
    def __init__(self, ctx):
        self.ctx = ctx

    def track(self, pair, prm, first):
        q = copy.copy(pair)
        if not first:
            q.img_ref = None                      # stream continuation: the previous current pyramid is on the device
        return self.ctx.track_batch([q], prm)[0]

    def carry(self, out, last_normal, pair):
        c = capi.CarryCase(None, None, None, last_normal, pair.K, pair.t_cur, pair.t_ref, W, H, n_keys=pair.n_keys)   # resident
        (n,) = self.ctx.set_predict_keypoints_and_mask([c])
        return c, n

    def detect(self, img, mask):
        return self.ctx.orb_cell_detect(img, 20, 7, mask=mask)


def run_loop(backend, frames, pairs):
    """returns per frame: (track outputs, survivors, new keypoints)"""
    prm = capi.default_params(pyramids=3)
    K = pairs[0].K
    keys = pairs[0].keys_ref_un[:N_WANT].copy()
    log = []
    for t in range(len(pairs)):
        p = copy.copy(pairs[t])
        p.keys_ref_un = np.ascontiguousarray(keys, np.float32)
        p.keys_ref = p.keys_ref_un
        out = backend.track(p, prm, first=(t == 0))
        last_normal = ((p.keys_ref_un - [K[0, 2], K[1, 2]]) * [np.float32(1.0 / K[0, 0]), np.float32(1.0 / K[1, 1])]).astype(np.float32)
        c, n = backend.carry(out, last_normal, p)
        cand_xy, cand_rs = backend.detect(frames[t + 1], c.mask)
        order = np.argsort(-cand_rs, kind="stable")[:max(0, N_WANT - n)]       # strongest first, ties in detection order
        new = cand_xy[order]
        log.append((out, c.keys_un[:n].copy(), c.index_in_last[:n].copy(), c.flow_velocity_last[:n].copy(), c.mask.copy(), new.copy()))
        keys = np.concatenate([c.keys_un[:n], new]).astype(np.float32)
    return log


def _compare(a, b):
    assert len(a) == len(b)
    for t, (x, y) in enumerate(zip(a, b)):
        helpers.assert_bit_exact(x[0], y[0], fields=set(helpers.FLOAT_FIELDS) | {"status", "pm_status"})
        assert x[0].n_predict == y[0].n_predict and x[0].n_iterations == y[0].n_iterations, t
        for k in range(1, 6):
            assert np.array_equal(x[k], y[k]), (t, k)
        assert len(x[1]) > 50, "the stream must keep most of its features"


@pytest.fixture(scope="module")
def stream():
    return synth.make_sequence(9300, N_FRAMES, width=W, height=H, n_keys=N_WANT, pyramids=3, border=24)


def test_frame_loop_restatement_against_the_reference_build(oracle, stream):
    from oracle import reference
    if reference.build() is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libpagk_ref.so is here")

    class RefBackend(OracleBackend):
        def detect(self, img, mask):
            xy, rs = self.m.orb_detect(img, 20, 7, mask=mask)
            o = np.lexsort((xy[:, 0], xy[:, 1]))                 # the octree scrambles the order of the (complete) set ...
            cx, cr = oracle.orb_cell_detect(img, 20, 7, mask=mask)
            oc = np.lexsort((cx[:, 0], cx[:, 1]))
            assert np.array_equal(xy[o], cx[oc]) and np.array_equal(rs[o], cr[oc])
            return cx, cr                                        # ... so hand the cell order on once the sets are equal

    frames, pairs = stream
    _compare(run_loop(RefBackend(reference), frames, pairs), run_loop(OracleBackend(oracle), frames, pairs))


@pytest.mark.gpu
def test_frame_loop_on_the_gpu(gpu_ctx, oracle, stream):
    frames, pairs = stream
    _compare(run_loop(GpuBackend(gpu_ctx), frames, pairs), run_loop(OracleBackend(oracle), frames, pairs))

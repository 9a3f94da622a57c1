"""Frame::SetPredictKeyPointsAndMask(): the restatement against the reference's own function
(src/frame.cpp compiled into the reference build), and the CUDA kernel against the restatement."""
import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth

FIELDS = ("keys", "keys_un", "keys_normal", "flow_velocity_last")


def make_case(seed, n=500, width=752, height=480, p_alive=0.8, want_mask=True):
    rng = np.random.default_rng(seed)
    K = synth.EUROC_K
    un = np.stack([rng.uniform(-3, width + 3, n), rng.uniform(-3, height + 3, n)], 1).astype(np.float32)   # some outside the image
    un[: n // 4] = np.round(un[: n // 4])
    pd = (un + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
    status = (rng.random(n) < p_alive).astype(np.uint8)
    last_normal = ((un + rng.normal(0, 3, (n, 2)) - [K[0, 2], K[1, 2]]) / [K[0, 0], K[1, 1]]).astype(np.float32)
    t_last = 1403715000.0 + seed
    return capi.CarryCase(pd, un, status, last_normal, K, t_last + 0.05 + 1e-4 * seed, t_last, width, height, want_mask=want_mask)


def _cases():
    return [make_case(1), make_case(2, n=1024, p_alive=1.0), make_case(3, n=37, p_alive=0.3), make_case(4, n=5, p_alive=0.0),
            make_case(5, n=300, width=640, height=480), make_case(6, n=2500, width=320, height=240, p_alive=0.6)]


def _same(a, b, n):
    for f in FIELDS:
        x, y = getattr(a, f)[:n], getattr(b, f)[:n]
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), f
    assert np.array_equal(a.index_in_last[:n], b.index_in_last[:n])
    if a.mask is not None and b.mask is not None:
        assert np.array_equal(a.mask, b.mask)


@pytest.fixture(scope="module")
def reference():
    from oracle import reference as r
    if r.build() is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libpagk_ref.so is here")
    r.load()
    return r


def test_restatement_matches_the_reference_function(reference, oracle):
    ca, cb = _cases(), _cases()
    rc, na = reference.set_predict_keypoints_and_mask(ca)
    rc2, nb = oracle.set_predict_keypoints_and_mask(cb)
    assert rc == 0 and rc2 == 0 and na == nb
    for a, b, n in zip(ca, cb, na):
        assert n == int(a.status.sum())
        _same(a, b, n)
        assert a.mask.min() == (0 if n else 1) and a.mask.max() == 1


@pytest.mark.gpu
def test_cuda_carry_over_bit_exact(gpu_ctx, oracle):
    cg, cc = _cases(), _cases()
    ng = gpu_ctx.set_predict_keypoints_and_mask(cg[:4])        # one mask size per call
    ng += gpu_ctx.set_predict_keypoints_and_mask(cg[4:5]) + gpu_ctx.set_predict_keypoints_and_mask(cg[5:6])
    rc, nc = oracle.set_predict_keypoints_and_mask(cc)
    assert rc == 0 and ng == nc
    for a, b, n in zip(cg, cc, nc):
        _same(a, b, n)


@pytest.mark.gpu
def test_cuda_carry_over_on_resident_results(gpu_ctx, oracle):
    """TrackFeatures() -> SetPredictKeyPointsAndMask() without moving the points off the device"""
    pairs = [synth.make_pair(9100 + i, width=320, height=240, n_keys=300, pyramids=3, border=20) for i in range(3)]
    prm = capi.default_params(pyramids=3)
    outs = gpu_ctx.track_batch(pairs, prm)
    cg, cc = [], []
    for p, o in zip(pairs, outs):
        K = p.K
        last_normal = ((p.keys_ref_un - [K[0, 2], K[1, 2]]) * [np.float32(1.0 / K[0, 0]), np.float32(1.0 / K[1, 1])]).astype(np.float32)
        cg.append(capi.CarryCase(None, None, None, last_normal, K, p.t_cur, p.t_ref, 320, 240, n_keys=p.n_keys))
        cc.append(capi.CarryCase(o.pt_predict, o.pt_predict_un, o.status, last_normal, K, p.t_cur, p.t_ref, 320, 240))
    ng = gpu_ctx.set_predict_keypoints_and_mask(cg)
    rc, nc = oracle.set_predict_keypoints_and_mask(cc)
    assert ng == nc and min(nc) > 100
    for a, b, n in zip(cg, cc, nc):
        _same(a, b, n)

"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _lcg_image(h, w, seed):
    x = (np.arange(h * w, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(seed * 1442695040888963407 + 1))
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    return ((x >> np.uint64(40)) & np.uint64(0xFF)).astype(np.uint8).reshape(h, w)


@pytest.mark.parametrize("shape,levels", [((480, 752), 4), ((480, 640), 3), ((1080, 1920), 4), ((2160, 3840), 5),
                                           ((135, 240), 3), ((77, 101), 3), ((270, 135), 2), ((64, 64), 5)])
def test_pyramid_bit_exact(gpu_ctx, oracle, shape, levels):
    h, w = shape
    if w > gpu_ctx.cfg.max_width or h > gpu_ctx.cfg.max_height:
        from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
        ctx = tracker.Context(max_width=w, max_height=h, max_keys=16, max_pairs=1, max_levels=levels, max_half_patch=5)
    else:
        ctx = gpu_ctx
    imgs = [_lcg_image(h, w, 1), _lcg_image(h, w, 2)]
    ctx.build_pyramids(imgs, levels)
    for k, img in enumerate(imgs):
        for l in range(levels):
            g = ctx.pyramid_level(k, l)
            c = oracle.pyramid_level(img, l)
            assert g.shape == c.shape
            assert np.array_equal(g, c), f"image {k} level {l}: {(g != c).sum()} bytes differ"
    if ctx is not gpu_ctx:
        ctx.close()


@pytest.mark.parametrize("e_type", [4, 2, 3, 5, 1])
def test_track_bit_exact_small(gpu_ctx, oracle, e_type):
    pairs = [synth.make_pair(7000 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(3)]
    prm = capi.default_params(e_type=e_type, pyramids=3)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 4)
    assert rc == 0
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c)


def test_track_config_B_pair(gpu_ctx, oracle):
    pairs = [synth.make_pair(2000 + i, **{k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}) for i in range(2)]
    prm = capi.default_params(pyramids=4)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 8)
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c)


@pytest.mark.parametrize("kernel", ["generic"])
def test_other_lk_kernels_still_bit_exact(cuda_lib, oracle, monkeypatch, kernel):
    """the any-patch-size kernel (PAGK_LK_KERNEL=generic) stays a second implementation of the same arithmetic to
    compare the production kernel against"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    monkeypatch.setenv("PAGK_LK_KERNEL", kernel)
    pairs = [synth.make_pair(7100 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(2)]
    prm = capi.default_params(pyramids=3)
    with tracker.Context(max_width=320, max_height=240, max_keys=200, max_pairs=2, max_levels=3) as ctx:
        gpu = ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 4)
    for g, c in zip(gpu, cpu):
        helpers.assert_bit_exact(g, c)


def test_border_features_bit_exact(gpu_ctx, oracle):
    """keypoints right up to the image border: clamped taps, guard row, gyro border rejections"""
    p = synth.make_pair(7200, width=320, height=240, n_keys=600, pyramids=3, border=0, sigma_w=2.0)
    prm = capi.default_params(pyramids=3)
    gpu = gpu_ctx.track_batch([p], prm)[0]
    rc, cpu = oracle.track(p, prm, 4)
    assert (cpu.status == 0).any() and (cpu.status == 1).any()
    helpers.assert_north_star(gpu, cpu)
    helpers.assert_bit_exact(gpu, cpu)


def _check(gpu_ctx, oracle, pairs, prm, threads=4):
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, threads)
    assert rc == 0
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c)
    return gpu, cpu


def test_flat_and_saturated_regions(gpu_ctx, oracle):
    """constant patches: the normal matrix is exactly singular, the LLT stops at pivot 0 and the update is NaN
    (reference src/patch_match.cpp:319-326) -> status 0, on both sides, feature for feature"""
    pairs = [synth.make_pair(7700 + i, width=320, height=240, n_keys=300, pyramids=3, border=12) for i in range(2)]
    for p in pairs:
        for img in (p.img_ref, p.img_cur):
            img[:, :110] = 128
            img[:80, 200:] = 255
            img[170:, 200:] = 0
    gpu, cpu = _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=3))
    assert any((c.pm_status == 0).any() for c in cpu) and any((c.pm_status == 1).any() for c in cpu)


@pytest.mark.parametrize("shape,levels", [((243, 331), 3), ((201, 177), 2), ((480, 640), 5)])
def test_odd_sizes_and_deep_pyramids(gpu_ctx, oracle, shape, levels):
    """odd level sizes go through OpenCV's fixed-point bilinear resize and the byte-wise window staging; five
    levels on 640 x 480 put the coarsest patches against the image border"""
    h, w = shape
    pairs = [synth.make_pair(7800 + i, width=w, height=h, n_keys=250, pyramids=levels, border=14) for i in range(2)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=levels))


def test_large_rotation_boxes_that_do_not_fit_the_window(gpu_ctx, oracle):
    """0.3 rad between the frames: strongly sheared patches whose sample box exceeds the staged window are sampled
    straight from the level by the cooperative pass; predictions that leave the image keep status 0"""
    pairs = [synth.make_pair(7900 + i, width=400, height=300, n_keys=300, pyramids=3, border=40, sigma_w=6.0) for i in range(3)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=3))


@pytest.mark.parametrize("levels,iterations", [(1, 1), (1, 10), (3, 1), (2, 30)])
def test_level_and_iteration_limits(gpu_ctx, oracle, levels, iterations):
    pairs = [synth.make_pair(8000, width=320, height=240, n_keys=200, pyramids=max(levels, 2), border=20)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=levels, iterations=iterations))


@pytest.mark.parametrize("n_pairs,n_keys", [(8, 20), (1, 1), (3, 700), (8, 1500)])
def test_batch_shapes_around_the_lane_cap(gpu_ctx, oracle, n_pairs, n_keys):
    """few features spread one per warp (cooperative pass only), many features fill every lane (lockstep pass)"""
    pairs = [synth.make_pair(8100 + i, width=320, height=240, n_keys=n_keys, pyramids=3, border=16) for i in range(n_pairs)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=3), threads=8)


def test_ragged_pairs_with_more_than_1024_features(gpu_ctx, oracle):
    """pairs above 1024 features go through the three-launch epilogue (elementwise phases over a grid of features, the
    ordered mean pixel error between them): counts that are no multiple of its block or chunk sizes, several chunks,
    and a pair below the limit in the same batch"""
    ns = [2500, 1025, 4100, 300]
    pairs = [synth.make_pair(8300 + i, width=640, height=480, n_keys=n, pyramids=3, border=16) for i, n in enumerate(ns)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=3), threads=8)


@pytest.mark.parametrize("half", [3, 7])
def test_patch_sizes_served_by_the_generic_kernel(gpu_ctx, oracle, half):
    """7 x 7 and 15 x 15 patches: no lane-kernel instantiation, the any-size kernel runs"""
    pairs = [synth.make_pair(8200 + half, width=320, height=240, n_keys=150, pyramids=3, border=24, half_patch=half)]
    _check(gpu_ctx, oracle, pairs, capi.default_params(pyramids=3, half_patch=half))


def test_frozen_golden_cases_on_gpu(gpu_ctx):
    """the committed golden fixtures (outputs of the reference build, tests/golden/lk_frozen.npz) against the CUDA path"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lk_frozen.npz"))
    for i in range(int(g["n"])):
        e_type = int(g[f"c{i}_in_e_type"])
        p = capi.PairInputs(g[f"c{i}_in_img_ref"], g[f"c{i}_in_img_cur"], g[f"c{i}_in_keys"], g[f"c{i}_in_imu_t"],
                            g[f"c{i}_in_imu_w"], float(g[f"c{i}_in_t_ref"]), float(g[f"c{i}_in_t_cur"]), g[f"c{i}_in_K"],
                            g[f"c{i}_in_Rbc"], dist=g[f"c{i}_in_dist"], n_dist=int(g[f"c{i}_in_n_dist"]))
        o = gpu_ctx.track_batch([p], capi.default_params(e_type=e_type, pyramids=int(g[f"c{i}_in_pyramids"]),
                                                         half_patch=int(g[f"c{i}_in_half_patch"])))[0]
        for name, arr in o.arrays().items():
            ref = g[f"c{i}_out_{name}"]
            if e_type == 6 and arr.dtype.kind == "f":   # device log() vs glibc log(): last-bit differences allowed
                assert np.allclose(arr, ref, rtol=0, atol=0.01, equal_nan=True), (i, name)
            elif e_type != 6:
                assert helpers.bits_equal(arr, ref).all(), (i, name)
        if e_type != 6:
            assert o.n_predict == int(g[f"c{i}_out_n_predict"]) and o.n_iterations == int(g[f"c{i}_out_n_iterations"])
        else:   # north-star tolerance for the mode with a transcendental in the loop
            assert (o.status == g[f"c{i}_out_status"]).mean() >= 0.999 or (o.status != g[f"c{i}_out_status"]).sum() <= 1


@pytest.mark.parametrize("e_type,levels", [(4, 3), (2, 3), (3, 3), (5, 3), (4, 4)])
def test_against_the_reference_build(gpu_ctx, e_type, levels):
    """the CUDA path against oracle/_ref/libpagk_ref.so (the reference's own sources compiled in the build container
    against stand-in third-party headers; the prebuilt library travels with the snapshot).  With 3 levels the call is
    GyroAidedTracker::TrackFeatures() itself."""
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref/libpagk_ref.so was not built (no /root/reference in the build container)")
    pairs = [synth.make_pair(7700 + 10 * e_type + i, width=320, height=240, n_keys=200, pyramids=levels, border=20)
             for i in range(2)]
    prm = capi.default_params(e_type=e_type, pyramids=levels)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, ref = reference.track_batch(pairs, prm, 4)
    assert rc == 0 and reference.last_path() == (0 if levels == 3 else 1)
    fields = set(helpers.FLOAT_FIELDS) | {"status", "pm_status"}   # per-feature pass counts are not observable there
    for g, c in zip(gpu, ref):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c, fields=fields)
        assert g.n_predict == c.n_predict and g.n_iterations == c.n_iterations
        assert helpers.bits_equal(g.Rcl, c.Rcl).all() and helpers.bits_equal(g.KRKinv, c.KRKinv).all()


def test_regularized_mode_within_tolerance(gpu_ctx, oracle):
    pairs = [synth.make_pair(7300 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(2)]
    prm = capi.default_params(e_type=6, pyramids=3)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 4)
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)


def test_ragged_and_empty_batches(gpu_ctx, oracle):
    """pairs with different keypoint counts in one batch, a pair with zero keypoints, and an empty batch"""
    ns = [0, 1, 33, 200]
    pairs = [synth.make_pair(7400 + i, width=320, height=240, n_keys=max(n, 1), pyramids=3, border=20) for i, n in enumerate(ns)]
    pairs[0].keys_ref_un = pairs[0].keys_ref_un[:0]; pairs[0].keys_ref = pairs[0].keys_ref[:0]
    prm = capi.default_params(pyramids=3)
    gpu = gpu_ctx.track_batch(pairs, prm)
    rc, cpu = oracle.track_batch(pairs, prm, 2)
    for g, c in zip(gpu, cpu):
        helpers.assert_bit_exact(g, c)
    assert gpu[0].n_predict == 0
    assert gpu_ctx.track_batch([], prm) == []


def test_distortion_and_single_homography(gpu_ctx, oracle):
    pairs = [synth.make_pair(7500, width=320, height=240, n_keys=150, pyramids=3, border=20, dist=synth.EUROC_DIST,
                             K=synth.scaled_euroc_K(320))]
    for method in (capi.PIXEL_AWARE_PREDICTION, capi.SINGLE_HOMOGRAPHY):
        prm = capi.default_params(pyramids=3, predict_method=method)
        g = gpu_ctx.track_batch(pairs, prm)[0]
        c = oracle.track(pairs[0], prm, 2)[1]
        helpers.assert_bit_exact(g, c)


@pytest.mark.parametrize("e_type", [3, 4])
def test_ncc_enabled_bit_exact(gpu_ctx, oracle, e_type):
    """bCalculateNCC_ = true (include/patch_match.h:49): PatchMatch::NCC, src/patch_match.cpp:433-469, with and
    without the affine warp; the tracker itself never switches it on, so this is the only place it runs"""
    pairs = [synth.make_pair(7600 + i, width=320, height=240, n_keys=200, pyramids=3, border=6) for i in range(2)]
    prm = capi.default_params(pyramids=3, e_type=e_type, calc_ncc=1)
    gpu = gpu_ctx.track_batch(pairs, prm)
    for g, p in zip(gpu, pairs):
        c = oracle.track(p, prm, 2)[1]
        helpers.assert_bit_exact(g, c)
        ok = c.pm_status == 1
        assert ok.any() and np.all(np.abs(g.ncc[ok]) <= 1.0 + 1e-6) and np.median(g.ncc[ok]) > 0.8


def test_stage_api_gyro_predict_and_patch_match(gpu_ctx, oracle):
    """the finer-grained entry points mirror GyroPredictFeatures and PatchMatch(...).OpticalFlowMultiLevel()"""
    p = synth.make_pair(7600, width=320, height=240, n_keys=180, pyramids=3, border=20)
    prm = capi.default_params(pyramids=3)
    g1 = gpu_ctx.gyro_predict(p, prm)
    c1 = oracle.gyro_predict(p, prm)
    helpers.assert_bit_exact(g1, c1, fields=["pt_predict_un", "pt_predict", "status", "affine", "corner_flows",
                                              "pt_corners_un", "pt_corners", "flows_predict_un"])
    for flags in (dict(consider_illumination=1, consider_affine_deformation=1), dict(consider_illumination=0, consider_affine_deformation=0),
                  dict(consider_illumination=1, consider_affine_deformation=0, pyramids=2, iterations=4)):
        s, keep = capi.patch_match_struct(p, c1.pt_predict_un, c1.status, c1.affine, regularization_penalty=0, **flags)
        g2 = gpu_ctx.patch_match(s, p.n_keys)
        rc, c2 = oracle.patch_match(s, p.n_keys, 2)
        assert rc == 0
        helpers.assert_bit_exact(g2, c2, fields=["pm_pt_un", "pm_pt", "pm_status", "pixel_error", "distance", "ncc", "iters"])


def test_reference_shaped_classes(gpu_ctx, oracle):
    """GyroAidedTracker / PatchMatch mirrors: same call sequence as Examples/Demo/RealSenseD435i.cpp:244-254"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    p = synth.make_pair(7700, width=320, height=240, n_keys=120, pyramids=3, border=20)
    cam = tracker.CameraParams(p.K, p.dist[:4], 320, 240)
    ref = tracker.Frame(p.t_ref, p.img_ref, p.keys_ref, p.keys_ref_un, None, cam)
    cur = tracker.Frame(p.t_cur, p.img_cur, mvImuFromLastFrame=(p.imu_t, p.imu_w), mpCameraParams=cam)
    Tbc = np.eye(4, dtype=np.float32); Tbc[:3, :3] = p.Rbc
    trk = tracker.GyroAidedTracker(gpu_ctx, ref, cur, Tbc, (0, 0, 0), None,
                                   tracker.GyroAidedTracker.GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                                   tracker.GyroAidedTracker.PIXEL_AWARE_PREDICTION, "", 5)
    n = trk.TrackFeatures()
    c = oracle.track(p, capi.default_params(pyramids=3), 2)[1]
    assert n == c.n_predict
    assert np.array_equal(trk.mvStatus, c.status) and helpers.bits_equal(trk.mvPtPredictUn, c.pt_predict_un).all()
    assert helpers.bits_equal(trk.GetRcl(), c.Rcl).all()
    trk.SetBackToFrame(cur)
    assert np.array_equal(cur.mvStatus, c.status) and cur.mvvFlowsPredictCorners.shape == (120, 4, 2)
    trk.SetType(tracker.GyroAidedTracker.OPENCV_OPTICAL_FLOW_PYR_LK)
    assert trk.TrackFeatures() == -1          # "Unsupport type!!! return -1;"
    # PatchMatch used directly, as GyroPredictFeaturesAndOpticalFlowRefined does (src/gyro_aided_tracker.cpp:280-283)
    trk2 = tracker.GyroAidedTracker(gpu_ctx, ref, cur, Tbc, type_=tracker.GyroAidedTracker.GYRO_PREDICT)
    trk2.TrackFeatures()
    pm = tracker.PatchMatch(trk2, 5, 10, 3, True, False, True, True, False)
    pm.OpticalFlowMultiLevel()
    assert np.array_equal(trk2.mvStatusAfterPatchMatched, c.pm_status)
    assert helpers.bits_equal(trk2.mvPtPredictAfterPatchMatchedUn, c.pm_pt_un).all()


def test_large_patch_generic_kernel(gpu_ctx, oracle):
    """21 x 21 patches (BASELINE config C's patch size): pagk_lk_lanes_kernel<10, *>, 28 slots per warp (the full-size config C
    and a 2048-feature case with level-granular items are in tests/test_baseline_configs.py)"""
    p = synth.make_pair(7800, width=480, height=360, n_keys=150, half_patch=10, pyramids=3, border=40,
                        K=synth.scaled_euroc_K(480))
    prm = capi.default_params(pyramids=3, half_patch=10)
    g = gpu_ctx.track_batch([p], prm)[0]
    c = oracle.track(p, prm, 4)[1]
    helpers.assert_north_star(g, c)
    helpers.assert_bit_exact(g, c)


def test_full_size_batch_properties(gpu_ctx, oracle):
    """BASELINE config B at full per-pair size (752x480, 1024 features, 4 levels), 8 pairs: size-independent
    properties (results do not depend on batching or on the pair's slot; pyramids idempotent) plus a spot check."""
    cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
    pairs = [synth.make_pair(2100 + i, **cfg) for i in range(8)]
    prm = capi.default_params(pyramids=4)
    a = gpu_ctx.track_batch(pairs, prm)
    b = gpu_ctx.track_batch(pairs[::-1], prm)[::-1]                    # other slots, other work order
    c = [gpu_ctx.track_batch([p], prm)[0] for p in pairs[:3]]          # one pair per call
    for x, y in zip(a, b):
        helpers.assert_bit_exact(x, y)
    for x, y in zip(a, c):
        helpers.assert_bit_exact(x, y)
    tot = sum(o.n_iterations for o in a)
    assert tot == sum(int(o.iters.sum()) for o in a) and tot > 8 * 1024 * 4
    ref = oracle.track(pairs[5], prm, 8)[1]
    helpers.assert_bit_exact(a[5], ref)


def test_level_granular_work_items_on_a_batch_larger_than_the_lanes(cuda_lib, oracle):
    """More features than the persistent kernel has lanes (148 SMs x 8 warps x 32): the launch queues one work item per
    feature LEVEL, and a level's result reaches the next level's lane through global memory.  Every output of all 40
    pairs bit-exact against the restatement, ragged key counts and skipped (status 0) features included."""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
    pairs = [synth.make_pair(2300 + i, **cfg) for i in range(40)]
    for p in pairs[::7]:                      # ragged: fewer keys than the batch maximum
        p.keys_ref_un = np.ascontiguousarray(p.keys_ref_un[:700])
        p.keys_ref = p.keys_ref_un
    pairs[3].keys_ref_un[::5] = (-50.0, -50.0)   # predictions outside the image: gyro status 0, skipped at every level
    prm = capi.default_params(pyramids=4)
    with tracker.Context(max_width=752, max_height=480, max_keys=1024, max_pairs=40, max_levels=4) as ctx:
        gpu = ctx.track_batch(pairs, prm)
        again = ctx.track_batch(pairs, prm)   # the progress words of the first launch must not satisfy the second
    rc, cpu = oracle.track_batch(pairs, prm, 8)
    assert rc == 0
    for g, a, c in zip(gpu, again, cpu):
        helpers.assert_bit_exact(g, c)
        helpers.assert_bit_exact(a, c)
    assert (gpu[3].status[::5] == 0).all()


def test_level_granular_work_items_forced_on_small_batches():
    """PAGK_LK_SPLIT=1 forces level-granular items whatever the batch size (lanes then mostly WAIT for the level above):
    the edge-case tests of this file rerun in a child process with it set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PAGK_LK_SPLIT="1")
    sel = ("track_bit_exact_small or border_features or flat_and_saturated or odd_sizes or large_rotation or "
           "level_and_iteration_limits or batch_shapes or ragged_and_empty or against_the_reference_build")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("share", [2, 3, 5])
def test_device_share_changes_only_the_launch_shape(cuda_lib, oracle, share):
    """pagk_set_device_share(h, n): a launch of the alignment kernel takes 1/n of every SM's CTA slots (the other handles of
    a pipeline run beside it); the results are the same bit for bit, also with two handles in flight at once"""
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    pairs = [synth.make_pair(8800 + i, width=320, height=240, n_keys=1024, pyramids=3, border=16) for i in range(17)]  # 17 408 features
    prm = capi.default_params(pyramids=3)
    rc, cpu = oracle.track_batch(pairs, prm, 8)
    assert rc == 0
    mk = lambda: tracker.Context(max_width=320, max_height=240, max_keys=1024, max_pairs=17, max_levels=3)
    with mk() as a, mk() as b:
        for c in (a, b):
            c.set_device_share(share)
            c.upload(pairs, prm)
        for _ in range(3):
            a.run(); b.run()
        outs = []
        for c in (a, b):
            c.synchronize()
            o = [capi.PairOutputs(p.n_keys) for p in pairs]
            c.download(o)
            outs.append(o)
    for o in outs:
        for g, c in zip(o, cpu):
            helpers.assert_bit_exact(g, c)
    with mk() as c:
        with pytest.raises(Exception):
            c.set_device_share(0)


def test_hand_over_tags_across_the_wrap_of_the_launch_counter():
    """the hand-over records of the lanes kernel are tagged launch number * 8 + levels finished; when the launch number is
    about to repeat (2^28 launches) the records are cleared and it restarts at 1.  A child process starts the counter 20
    launches below the wrap (PAGK_DEBUG_LK_EPOCH) with level-granular items forced and reruns one batch 48 times: every run
    bit-identical to the oracle."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np\n"
        "from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker\n"
        "from oracle import oracle\n"
        "from tests import helpers\n"
        "pairs = [synth.make_pair(8700 + i, width=320, height=240, n_keys=300, pyramids=3, border=16) for i in range(3)]\n"
        "prm = capi.default_params(pyramids=3)\n"
        "rc, cpu = oracle.track_batch(pairs, prm, 4)\n"
        "assert rc == 0\n"
        "with tracker.Context(max_width=320, max_height=240, max_keys=300, max_pairs=3, max_levels=3) as ctx:\n"
        "    for it in range(48):\n"
        "        gpu = ctx.track_batch(pairs, prm)\n"
        "        for g, c in zip(gpu, cpu):\n"
        "            helpers.assert_bit_exact(g, c)\n"
        "print('wrap ok')\n")
    env = dict(os.environ, PAGK_LK_SPLIT="1", PAGK_DEBUG_LK_EPOCH=str(0x0fffffff - 20))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "wrap ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_stage_timing_off_changes_only_the_clocks(gpu_ctx):
    """pagk_set_stage_timing(h, 0): no CUDA event between the kernels of a run; results identical, the whole device time is
    booked on the patch alignment"""
    pairs = [synth.make_pair(8600 + i, width=320, height=240, n_keys=200, pyramids=3, border=20) for i in range(2)]
    prm = capi.default_params(pyramids=3)
    on = gpu_ctx.track_batch(pairs, prm)
    t_on = gpu_ctx.last_run_ms()
    gpu_ctx.set_stage_timing(False)
    try:
        off = gpu_ctx.track_batch(pairs, prm)
        t_off = gpu_ctx.last_run_ms()
    finally:
        gpu_ctx.set_stage_timing(True)
    for a, b in zip(on, off):
        helpers.assert_bit_exact(a, b)
    assert t_on["pyramid"] > 0 and t_on["lk"] > 0 and t_on["filter"] > 0
    assert t_off["pyramid"] == 0 and t_off["predict"] == 0 and t_off["filter"] == 0 and t_off["lk"] == t_off["total"] > 0
    assert off[0].struct.t_gyro_predict == 0 and off[0].struct.t_opt_flow > 0


def test_stream_continuation_reuses_the_previous_current_pyramid(gpu_ctx, oracle):
    """img_ref = NULL on every pair: frame t is tracked against frame t-1 of the same stream, whose pyramid stays on the device
    (only the new image is uploaded, only its pyramid is built).  Three streams, five frames: every step bit-exact against the
    restatement run on the explicit pair, and against the pairwise call."""
    import copy
    seqs = [synth.make_sequence(8800 + s, 5, width=320, height=240, n_keys=200, pyramids=3, border=20) for s in range(3)]
    prm = capi.default_params(pyramids=3)
    for t in range(4):
        pairs = [seq[1][t] for seq in seqs]
        if t == 0:
            got = gpu_ctx.track_batch(pairs, prm)                 # the first pair of a stream brings both images
        else:
            cont = []
            for p in pairs:
                q = copy.copy(p)
                q.img_ref = None
                cont.append(q)
            got = gpu_ctx.track_batch(cont, prm)
        rc, cpu = oracle.track_batch(pairs, prm, 4)
        assert rc == 0
        for g, c in zip(got, cpu):
            helpers.assert_bit_exact(g, c)
        gpu_ctx._pyr_wh = (320, 240)
        for lvl in range(3):                                        # the pyramids on the device are those of this step's pair
            assert np.array_equal(gpu_ctx.pyramid_level(0, lvl), oracle.pyramid_level(pairs[0].img_ref, lvl))
            assert np.array_equal(gpu_ctx.pyramid_level(5, lvl), oracle.pyramid_level(pairs[2].img_cur, lvl))


def test_stream_continuation_needs_a_previous_batch(cuda_lib):
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    import copy
    p = synth.make_pair(8900, width=160, height=120, n_keys=20, pyramids=3, border=12, margin=32)
    q = copy.copy(p)
    q.img_ref = None
    prm = capi.default_params(pyramids=3)
    with tracker.Context(max_width=160, max_height=120, max_keys=32, max_pairs=2, max_levels=3) as ctx:
        with pytest.raises(tracker.PagkError):
            ctx.track_batch([q], prm)                              # nothing to continue from
        ctx.track_batch([p], prm)
        ctx.track_batch([q], prm)                                  # fine now
        with pytest.raises(tracker.PagkError):
            ctx.track_batch([q, q], prm)                           # more streams than the previous batch had
        with pytest.raises(tracker.PagkError):
            ctx.track_batch([p, q], prm)                           # mixed


@pytest.mark.gpu
@pytest.mark.parametrize("n_imu", [0, 1, 2, 3])
def test_short_imu_vectors(gpu_ctx, oracle, n_imu):
    """no / one / two / three gyro samples (src/gyro_aided_tracker.cpp:521-562: size() - 1 intervals)"""
    import copy
    p = synth.make_pair(8100 + n_imu, width=320, height=240, n_keys=64, pyramids=3, border=24)
    q = copy.copy(p)
    q.imu_t, q.imu_w = p.imu_t[:n_imu].copy(), p.imu_w[:n_imu].copy()
    prm = capi.default_params(pyramids=3)
    (g,) = gpu_ctx.track_batch([q], prm)
    rc, c = oracle.track(q, prm, 1)
    assert rc == 0
    helpers.assert_bit_exact(g, c)


@pytest.mark.gpu
def test_images_with_a_row_pitch(gpu_ctx, oracle):
    """a cv::Mat that is a region of a larger one has step > cols (GetPixelValue indexes with img.step,
    src/patch_match.cpp:397-403): pagk_pair_in.pitch carries it; the result is that of the packed image"""
    import ctypes as C
    p = synth.make_pair(8300, width=320, height=240, n_keys=100, pyramids=3, border=24)
    prm = capi.default_params(pyramids=3)
    rc, c = oracle.track(p, prm, 1)
    assert rc == 0
    pad = 40
    big_ref = np.full((240, 320 + pad), 255, np.uint8); big_ref[:, :320] = p.img_ref
    big_cur = np.full((240, 320 + pad), 7, np.uint8); big_cur[:, :320] = p.img_cur
    ins = capi.make_in_array([p])
    ins[0].img_ref, ins[0].img_cur = big_ref.ctypes.data_as(capi._u8p), big_cur.ctypes.data_as(capi._u8p)
    ins[0].pitch = big_ref.strides[0]
    out = capi.PairOutputs(p.n_keys)
    oarr = capi.make_out_array([out])
    assert gpu_ctx.lib.pagk_track_batch(gpu_ctx.handle, C.byref(prm), 1, ins, oarr) == capi.PAGK_OK
    capi.sync_out_array(oarr, [out])
    helpers.assert_bit_exact(out, c)

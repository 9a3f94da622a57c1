"""The drivers' input side: settings / image list / IMU log / keypoint-file parsers against files in
the reference's formats (Examples/Demo/RealSenseD435i.cpp:74-141, 168-182, 207-217; include/common.h:49-114;
src/frame.cpp:222-262), and the frame loop over a recorded sequence: CUDA path against the oracle, bit for bit."""
import os

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, sequence, synth
from tests import helpers

W, H, N_FRAMES, N_WANT = 320, 240, 5, 120

SETTINGS = """%YAML:1.0

dataset: "PKUSZ_RealSenseD435i_sequence"
datasetDir: "data/seq"
outputFile: "/output/"
downSampleRate: 1  # useless
KeyPointNumber: {n}
ThresholdOfPredictNewKeyPoint: 1.0
HalfPatchSize: 5
LoadDetectedKeypoints: {load}
DetectedKeypointsFile: "/data/SuperPoints/seq"
Camera.type: "PinHole"
Camera.fx: {fx!r}
Camera.fy: {fy!r}
Camera.cx: {cx!r}
Camera.cy: {cy!r}

Camera.k1: 0.0
Camera.k2: 0.0
Camera.p1: 0.0
Camera.p2: 0.0

Camera.width: {w}
Camera.height: {h}
Camera.fps: 20
Camera.RGB: 1

# Transformation from camera to body-frame (imu)
Tbc:
    [{tbc}]

IMU.NoiseGyro: 0.005       # rad/s/sqrt(Hz)
IMU.NoiseAcc: 0.1
IMU.GyroWalk: 0.0002
IMU.AccWalk: 0.005
IMU.Frequency: 200
"""


def write_sequence(root, frames, pairs, with_keypoints=False):
    """a dataset directory in the layout of the reference's demo: <root>/seq/{image_file_list.txt, imu.txt, cam0/<ns>.png}"""
    import cv2
    d = os.path.join(root, "seq")
    os.makedirs(os.path.join(d, "cam0"))
    times = [pairs[0].t_ref] + [p.t_cur for p in pairs]
    names = []
    for img, t in zip(frames, times):
        ns = int(round(t * 1e9))
        names.append(ns)
        assert cv2.imwrite(os.path.join(d, "cam0", f"{ns}.png"), img)
    with open(os.path.join(d, "image_file_list.txt"), "w") as f:
        f.write("".join(f"/cam0/{ns}.png\n" for ns in names))
    rows = {}
    for p in pairs:                                   # the per-pair windows of the generator overlap: one merged log
        for t, w in zip(p.imu_t, p.imu_w):
            rows[int(round(t * 1e9))] = w
    with open(os.path.join(d, "imu.txt"), "w") as f:
        for ns in sorted(rows):
            w = rows[ns]
            f.write(f"{ns} 0.0 9.81 0.0 {float(w[0])!r} {float(w[1])!r} {float(w[2])!r}\n")
    K, Rbc = pairs[0].K, pairs[0].Rbc
    Tbc = np.eye(4, dtype=np.float64)
    Tbc[:3, :3] = Rbc
    Tbc[:3, 3] = [-0.0055, 0.0051, 0.01174]
    with open(os.path.join(root, "settings.yaml"), "w") as f:
        f.write(SETTINGS.format(n=N_WANT, load=int(with_keypoints), fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]),
                                cy=float(K[1, 2]), w=frames[0].shape[1], h=frames[0].shape[0],
                                tbc=",\n    ".join(", ".join(repr(float(x)) for x in row) for row in Tbc)))
    if with_keypoints:
        kd = os.path.join(root, "SuperPoints", "seq")
        os.makedirs(kd)
        rng = np.random.default_rng(5)
        with open(os.path.join(kd, "corresponds.txt"), "w") as f:
            for ns in names:
                f.write(f"{ns * 1e-9!r}, {ns}\n")
        for ns in names:
            pts = np.stack([rng.integers(24, frames[0].shape[1] - 24, 400), rng.integers(24, frames[0].shape[0] - 24, 400)], 1)
            with open(os.path.join(kd, f"{ns}.txt"), "w") as f:
                f.write("".join(f"{i}, {float(x)}, {float(y)}\n" for i, (x, y) in enumerate(pts)))
    return d, names


@pytest.fixture(scope="module")
def recorded(tmp_path_factory):
    frames, pairs = synth.make_sequence(9400, N_FRAMES, width=W, height=H, n_keys=N_WANT, pyramids=3, border=24)
    root = str(tmp_path_factory.mktemp("pagk_seq"))
    d, names = write_sequence(root, frames, pairs, with_keypoints=True)
    return dict(root=root, dir=d, names=names, frames=frames, pairs=pairs)


def test_settings_file(recorded):
    s = sequence.load_configure_file(os.path.join(recorded["root"], "settings.yaml"))
    p = recorded["pairs"][0]
    assert np.array_equal(s.K, p.K) and s.K.dtype == np.float32
    assert np.array_equal(s.Rbc, p.Rbc)
    assert s.dist.shape == (4,) and not s.dist.any()             # no Camera.k3 -> four coefficients (n_dist = 4)
    assert (s.width, s.height, s.fps) == (W, H, 20)
    assert s.keypoint_number == N_WANT and s.half_patch_size == 5 and s.load_detected_keypoints
    assert s.dataset == "PKUSZ_RealSenseD435i_sequence" and s.dataset_dir == "data/seq" and s.imu_frequency == 200
    with pytest.raises(KeyError):
        bad = os.path.join(recorded["root"], "bad.yaml")
        open(bad, "w").write("%YAML:1.0\nKeyPointNumber: 3\n")
        sequence.load_configure_file(bad)


def test_reference_settings_file_when_present():
    path = "/root/reference/Examples/Demo/RealSenseD435i.yaml"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this machine")
    s = sequence.load_configure_file(path)
    assert s.keypoint_number == 500 and s.half_patch_size == 5 and (s.width, s.height, s.fps) == (640, 480, 15)
    assert s.K[0, 0] == np.float32(394.5643528049837) and s.dist[0] == np.float32(-0.0027697209770466296) and s.dist.size == 4
    assert np.array_equal(s.Rbc, np.eye(3, dtype=np.float32)) and not s.load_detected_keypoints


def test_image_list_and_imu_log(recorded):
    imgs = sequence.read_image_file_list(recorded["dir"])
    assert len(imgs) == N_FRAMES
    for (t, path), ns in zip(imgs, recorded["names"]):
        assert t == ns * 1e-9 and path == recorded["dir"] + f"/cam0/{ns}.png" and os.path.exists(path)
    log = sequence.read_imu_txt(os.path.join(recorded["dir"], "imu.txt"))
    assert log.t.dtype == np.float64 and log.w.dtype == np.float32 and log.a.shape == log.w.shape == (log.t.size, 3)
    assert np.all(np.diff(log.t) > 0) and np.all(log.a[:, 1] == np.float32(9.81))
    assert sequence._stol_ns("  1627889784040685824.png") == 1627889784040685824 * 1e-9     # stol stops at the first non-digit
    with pytest.raises(ValueError):
        sequence._stol_ns("abc")


def test_imu_windows_are_the_half_open_frame_intervals():
    """Examples/Demo/RealSenseD435i.cpp:207-217: skip samples older than time_prev, hand out those older than time_cur"""
    t = 100.0 + np.arange(60) * 0.005
    log = sequence.ImuLog(t, np.zeros((60, 3), np.float32), np.arange(180, dtype=np.float32).reshape(60, 3))
    feed = sequence.ImuFeed(log)
    frames = [100.012, 100.0625, 100.11, 100.11 + 1e-9, 100.2, 100.5, 100.7]
    prev = 0.0
    for k, cur in enumerate(frames):
        idx = feed.window(prev, cur)
        if k == 0:
            assert idx.size == 0                                   # first frame: time_prev == 0
        elif cur <= t[-1]:
            assert np.array_equal(idx, np.nonzero((t >= prev) & (t < cur))[0]), k
        elif k == 5:
            assert np.array_equal(idx, np.nonzero(t >= prev)[0]) and not feed.valid_imu     # the log runs out: last sample once
        else:
            assert idx.size == 0                                   # ... and never again
        prev = cur


def test_keypoint_files(recorded):
    kd = os.path.join(recorded["root"], "SuperPoints", "seq")
    table = sequence.read_time_correspondences(os.path.join(kd, "corresponds.txt"))
    assert [s for _, s in table] == [str(ns) for ns in recorded["names"]]
    ns = recorded["names"][2]
    assert sequence.find_time_correspondence_index(table, ns * 1e-9 + 5e-5) == 2
    assert sequence.find_time_correspondence_index(table, ns * 1e-9 + 0.02) == -1
    pts = sequence.read_detected_keypoints(os.path.join(kd, f"{ns}.txt"))
    assert pts.shape == (400, 2) and pts.dtype == np.float32 and np.all(pts == np.floor(pts))
    mask = np.full((H, W), 255, np.uint8)
    mask[:, : W // 2] = 0
    sel = sequence.filter_new_keypoints(pts, mask, 50)
    want = pts[pts[:, 0] >= W // 2][:50]
    assert np.array_equal(sel, want)
    assert sequence.filter_new_keypoints(pts, mask, 0).shape == (0, 2)
    assert np.array_equal(sequence.filter_new_keypoints(pts, None, 7), pts[:7])


def test_reference_superpoint_files_when_present():
    z = "/root/reference/Examples/Demo/data/PKUSZ_RealSenseD435i_sequence/SuperPoints.zip"
    if not os.path.exists(z):
        pytest.skip("the reference tree is not on this machine")
    import tempfile, zipfile
    with tempfile.TemporaryDirectory() as d:
        zipfile.ZipFile(z).extractall(d)
        kd = os.path.join(d, "SuperPoints", "sequence_1")
        table = sequence.read_time_correspondences(os.path.join(kd, "corresponds.txt"))
        assert len(table) == 149 and table[0] == (1627889784.040686, "1627889784040685824")
        for i in (0, 77, 148):                                    # the table's seconds match the stems' nanoseconds
            assert sequence.find_time_correspondence_index(table, int(table[i][1]) * 1e-9) == i
        pts = sequence.read_detected_keypoints(os.path.join(kd, table[0][1] + ".txt"))
        assert pts.shape[1] == 2 and pts.shape[0] > 100 and np.array_equal(pts[0], [455.0, 44.0])
        assert pts[:, 0].max() < 640 and pts[:, 1].max() < 480


class OracleContext:
    """the three calls `sequence.replay` makes, answered by the CPU oracle (test infrastructure)"""

    def __init__(self, oracle):
        self.o, self.prev_cur, self.last = oracle, None, None

    def track_batch(self, pairs, prm):
        (p,) = pairs
        if p.img_ref is None:                                     # stream continuation
            p = capi.PairInputs(self.prev_cur, p.img_cur, p.keys_ref_un, p.imu_t, p.imu_w, p.t_ref, p.t_cur, p.K, p.Rbc,
                                dist=p.dist, n_dist=p.n_dist)
        rc, out = self.o.track(p, prm, 2)
        assert rc == 0
        self.prev_cur, self.last = p.img_cur, out
        return [out]

    def set_predict_keypoints_and_mask(self, cases):
        (c,) = cases                                              # "resident" inputs: the results of the run before
        c.pt_predict, c.pt_predict_un, c.status = self.last.pt_predict, self.last.pt_predict_un, self.last.status
        rc, n = self.o.set_predict_keypoints_and_mask([c])
        assert rc == 0
        return n

    def orb_cell_detect(self, img, ini_th=20, min_th=7, mask=None):
        return self.o.orb_cell_detect(img, ini_th, min_th, mask=mask)


def _run(ctx, recorded, keypoint_dir=None):
    s = sequence.load_configure_file(os.path.join(recorded["root"], "settings.yaml"))
    seq = sequence.RecordedSequence(recorded["dir"])
    return list(sequence.replay(ctx, seq, s, keypoint_dir=keypoint_dir))


def _check_loop(res, recorded):
    assert len(res) == N_FRAMES and res[0].outputs is None and res[0].n_new == N_WANT
    for k, r in enumerate(res[1:], 1):
        assert r.t == recorded["names"][k] * 1e-9
        assert r.n_ref == res[k - 1].keys_un.shape[0] == N_WANT
        assert r.n_carried > N_WANT // 2, "the stream must keep most of its features"
        assert r.n_carried + r.n_new == r.keys_un.shape[0] <= N_WANT
        assert np.all(r.index_in_last[: r.n_carried] >= 0) and np.all(r.index_in_last[r.n_carried:] == -1)
        st = r.outputs.status.astype(bool)
        assert r.n_carried <= int(st.sum())


def test_replay_decodes_what_was_recorded(recorded):
    seq = sequence.RecordedSequence(recorded["dir"])
    fr = list(seq.frames())
    assert len(seq) == len(fr) == N_FRAMES and fr[0].imu_t.size == 0
    for k, f in enumerate(fr):
        assert np.array_equal(sequence.read_gray(f.path), recorded["frames"][k])      # gray PNG -> BGR -> RGB2GRAY is lossless
        if k:
            assert f.imu_t.size >= 9 and f.imu_t[0] >= fr[k - 1].t and f.imu_t[-1] < f.t and f.imu_w.shape == (f.imu_t.size, 3)


def test_replay_with_the_oracle(oracle, recorded):
    _check_loop(_run(OracleContext(oracle), recorded), recorded)
    res = _run(OracleContext(oracle), recorded, keypoint_dir=os.path.join(recorded["root"], "SuperPoints", "seq"))
    _check_loop(res, recorded)
    assert np.all(res[0].keys_un == np.floor(res[0].keys_un))     # keypoints from the files are whole pixels


@pytest.mark.gpu
def test_replay_on_the_gpu(gpu_ctx, oracle, recorded):
    for kd in (None, os.path.join(recorded["root"], "SuperPoints", "seq")):
        a, b = _run(gpu_ctx, recorded, kd), _run(OracleContext(oracle), recorded, kd)
        _check_loop(a, recorded)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert (x.n_ref, x.n_predict, x.n_carried, x.n_new) == (y.n_ref, y.n_predict, y.n_carried, y.n_new)
            assert np.array_equal(x.keys_un, y.keys_un) and np.array_equal(x.index_in_last, y.index_in_last)
            if x.outputs is not None:
                helpers.assert_bit_exact(x.outputs, y.outputs)


def test_cpp_readers_agree_with_the_python_mirror(recorded, tmp_path):
    """include/pagk_sequence.hpp (the C++ host side) against sequence.py on the same files, value for value"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "sequence_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "sequence_dump.cpp"), "-o", exe])
    kd = os.path.join(recorded["root"], "SuperPoints", "seq")
    r = subprocess.run([exe, os.path.join(recorded["root"], "settings.yaml"), recorded["dir"], kd], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    s = sequence.load_configure_file(os.path.join(recorded["root"], "settings.yaml"))
    f = lines[0].split(" ", 17)
    assert f[0] == "settings"
    assert [float.fromhex(x) for x in f[1:9]] == [float(v) for v in (s.K[0, 0], s.K[1, 1], s.K[0, 2], s.K[1, 2], *s.dist[:4])]
    assert [int(x) for x in f[9:13]] == [0, s.width, s.height, s.fps]
    assert float.fromhex(f[13]) == s.threshold_of_predict_new_keypoint
    assert [int(x) for x in f[14:17]] == [s.keypoint_number, s.half_patch_size, int(s.load_detected_keypoints)]
    assert f[17] == f"{s.dataset}|{s.dataset_dir}|{s.detected_keypoints_file}"
    assert [float.fromhex(x) for x in lines[1].split()[1:]] == [float(v) for v in s.Tbc.reshape(-1)]
    seq = sequence.RecordedSequence(recorded["dir"])
    table = sequence.read_time_correspondences(os.path.join(kd, "corresponds.txt"))
    it = iter(lines[2:])
    for fr in seq.frames():
        head = next(it).split()
        idx = sequence.find_time_correspondence_index(table, fr.t)
        assert head[0] == "frame" and float.fromhex(head[1]) == fr.t and head[2:] == [fr.path, str(fr.imu_t.size), str(idx)]
        for t, w in zip(fr.imu_t, fr.imu_w):
            m = next(it).split()
            assert m[0] == "imu" and float.fromhex(m[1]) == t and [float.fromhex(x) for x in m[5:8]] == [float(v) for v in w]
            assert [float.fromhex(x) for x in m[2:5]] == [0.0, float(np.float32(9.81)), 0.0]
        pts = sequence.read_detected_keypoints(os.path.join(kd, table[idx][1] + ".txt"))
        k = next(it).split()
        assert k[0] == "keypoints" and int(k[1]) == pts.shape[0]
        assert float.fromhex(k[2]) == float(pts[:, 0].astype(np.float64).sum()) and float.fromhex(k[3]) == float(pts[:, 1].astype(np.float64).sum())
    assert next(it, None) is None

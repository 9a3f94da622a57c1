"""The drivers' input side: the files the reference's demo reads and the frame loop around the
tracker, so that a recorded sequence replays through the CUDA path end to end.

Reference (all under /root/reference):
  * settings file                 include/common.h:49-103  (`loadConfigureFile`, an OpenCV FileStorage YAML 1.0 file)
  * image_file_list.txt           Examples/Demo/RealSenseD435i.cpp:74-100  (`getNextFrame`)
  * imu.txt                       Examples/Demo/RealSenseD435i.cpp:102-141 (`getNextIMU`)
  * IMU samples of a frame pair   Examples/Demo/RealSenseD435i.cpp:207-217
  * corresponds.txt               Examples/Demo/RealSenseD435i.cpp:168-182, include/common.h:105-114
  * SuperPoint keypoint files     src/frame.cpp:222-240 (`Frame::LoadDetectedKeypointFromFile`)
  * the frame loop                Examples/Demo/RealSenseD435i.cpp:198-296

Parsing is host work and stays on the host.  Everything between two parses -- rectification, pyramids, prediction, patch
alignment, filter, carry-over with the occupancy mask, per-cell FAST top-up -- runs on the device through `tracker.Context`
(`replay`); nothing here computes a tracking result on the CPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import capi


# ------------------------------------------------------------------------------------------------
# settings (include/common.h:49-103)
# ------------------------------------------------------------------------------------------------
@dataclass
class Settings:
    """What `loadConfigureFile` leaves in the demo's globals, plus the three keys `main` reads itself."""
    K: np.ndarray                      # 3x3 f32 from Camera.fx, fy, cx, cy
    dist: np.ndarray                   # k1 k2 p1 p2 (k3): 4 values unless Camera.k3 is a real number
    width: int
    height: int
    fps: int
    Tbc: np.ndarray                    # 4x4 f32
    keypoint_number: int
    threshold_of_predict_new_keypoint: float
    half_patch_size: int
    load_detected_keypoints: bool = False
    detected_keypoints_file: str = ""
    dataset: str = ""
    dataset_dir: str = ""
    imu_frequency: int = 200
    raw: dict = field(default_factory=dict)

    @property
    def Rbc(self) -> np.ndarray:
        return np.ascontiguousarray(self.Tbc[:3, :3], np.float32)


def load_configure_file(path: str) -> Settings:
    """`loadConfigureFile` (include/common.h:49-103) for the YAML subset the reference's settings files use: scalars and
    flow sequences.  cv::FileStorage writes the directive as `%YAML:1.0`, which YAML proper spells `%YAML 1.0`."""
    import yaml
    text = open(path, "r").read()
    if text.lstrip().startswith("%YAML:"):
        head, _, rest = text.partition("\n")
        text = "%YAML 1.1\n---\n" + rest
    d = yaml.safe_load(text)
    if not isinstance(d, dict):
        raise ValueError(f"{path}: not a settings file")

    def need(key):
        if key not in d:
            raise KeyError(f"{path}: missing '{key}'")
        return d[key]

    f32 = np.float32
    fx, fy, cx, cy = (f32(need("Camera." + k)) for k in ("fx", "fy", "cx", "cy"))      # `float fx = fSettings[...]`
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    dist = [f32(need("Camera." + k)) for k in ("k1", "k2", "p1", "p2")]
    k3 = d.get("Camera.k3")
    if isinstance(k3, float):                                                           # `node.isReal()`
        dist.append(f32(k3))
    tbc = np.asarray(need("Tbc"), np.float32).reshape(-1)
    if tbc.size != 16:
        raise ValueError(f"{path}: Tbc needs 16 values, has {tbc.size}")
    return Settings(
        K=K, dist=np.asarray(dist, np.float32), width=int(need("Camera.width")), height=int(need("Camera.height")),
        fps=int(need("Camera.fps")), Tbc=tbc.reshape(4, 4),
        keypoint_number=int(need("KeyPointNumber")),
        threshold_of_predict_new_keypoint=float(need("ThresholdOfPredictNewKeyPoint")),
        half_patch_size=int(d.get("HalfPatchSize", 5)),
        load_detected_keypoints=int(d.get("LoadDetectedKeypoints", 0)) == 1,
        detected_keypoints_file=str(d.get("DetectedKeypointsFile", "")),
        dataset=str(d.get("dataset", "")), dataset_dir=str(d.get("datasetDir", "")),
        imu_frequency=int(d.get("IMU.Frequency", 200)), raw=d)


# ------------------------------------------------------------------------------------------------
# image list and IMU log
# ------------------------------------------------------------------------------------------------
def _stol_ns(s: str) -> float:
    """`std::stol(s) * 1e-9`: leading whitespace, optional sign, the longest run of digits; the product in double."""
    s = s.lstrip()
    i = 1 if s[:1] in "+-" else 0
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    if j == i:
        raise ValueError(f"stol: no digits in {s!r}")
    return int(s[:j]) * 1e-9


def read_image_file_list(dataset_dir: str) -> List[Tuple[float, str]]:
    """`getNextFrame` (Examples/Demo/RealSenseD435i.cpp:74-100): every line of <dataset_dir>/image_file_list.txt is a path
    that is appended to dataset_dir as it stands; the timestamp is the file name between the last '/' and ".png", in
    nanoseconds."""
    out = []
    with open(os.path.join(dataset_dir, "image_file_list.txt"), "r") as f:
        for line in f.read().split("\n"):
            if line == "":
                continue
            line = line.rstrip("\r")
            pos1, pos2 = line.rfind("/"), line.rfind(".png")
            if pos2 < 0:
                raise ValueError(f"image_file_list.txt: no '.png' in {line!r}")
            out.append((_stol_ns(line[pos1 + 1:pos2]), dataset_dir + line))
    return out


@dataclass
class ImuLog:
    t: np.ndarray   # [n] float64 seconds
    a: np.ndarray   # [n][3] float32 (IMU::Point stores cv::Point3f)
    w: np.ndarray   # [n][3] float32


def read_imu_txt(path: str) -> ImuLog:
    """`getNextIMU` (Examples/Demo/RealSenseD435i.cpp:102-141): whitespace-separated `t_ns ax ay az wx wy wz` per line;
    values are parsed as double and stored as float."""
    t, a, w = [], [], []
    with open(path, "r") as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if len(p) < 7:
                raise ValueError(f"{path}: expected 7 fields, got {len(p)}: {line!r}")
            t.append(_stol_ns(p[0]))
            a.append([float(x) for x in p[1:4]])
            w.append([float(x) for x in p[4:7]])
    return ImuLog(np.asarray(t, np.float64), np.asarray(a, np.float64).astype(np.float32).reshape(-1, 3),
                  np.asarray(w, np.float64).astype(np.float32).reshape(-1, 3))


class ImuFeed:
    """The demo's IMU cursor (Examples/Demo/RealSenseD435i.cpp:196-217): one sample is always read ahead (`last_imu`);
    for a frame pair the cursor first skips samples older than time_prev - delay, then hands out every sample older than
    time_cur - delay.  The sample that ends a window stays in `last_imu` and opens the next one.  When the log runs out
    `valid_imu` goes false and stays false, and the last sample is handed out once, as in the reference."""

    def __init__(self, log: ImuLog):
        self.log = log
        self._next = 0
        self._last = None            # index of `last_imu`
        self.valid_imu = True
        self._get_next()             # `IMU::Point last_imu; getNextIMU(last_imu);`

    def _get_next(self) -> bool:
        if self._next < self.log.t.shape[0]:
            self._last = self._next
            self._next += 1
            return True
        return False

    def window(self, time_prev: float, time_cur: float, delay: float = 0.0) -> np.ndarray:
        """indices into the log of `vImuMeas` for the pair (time_prev, time_cur); empty for the first frame (time_prev == 0)"""
        idx: List[int] = []
        if time_prev != 0 and self._last is not None:
            while self.log.t[self._last] < time_prev - delay and self._get_next():
                pass
            while self.log.t[self._last] < time_cur - delay and self.valid_imu:
                idx.append(self._last)
                self.valid_imu = self._get_next()
        return np.asarray(idx, np.int64)


# ------------------------------------------------------------------------------------------------
# detected-keypoint files
# ------------------------------------------------------------------------------------------------
def read_time_correspondences(path: str) -> List[Tuple[float, str]]:
    """corresponds.txt (Examples/Demo/RealSenseD435i.cpp:168-182): `<seconds>, <file stem>` per line; the stem starts two
    characters behind the comma."""
    out = []
    with open(path, "r") as f:
        for line in f.read().split("\n"):
            if line == "":
                continue
            p = line.find(",")
            if p < 0:
                raise ValueError(f"{path}: no ',' in {line!r}")
            out.append((float(line[:p]), line[p + 2:].rstrip("\r")))
    return out


def find_time_correspondence_index(table: Sequence[Tuple[float, str]], t: float) -> int:
    """`findTimeCorrespondenIndex` (include/common.h:105-114): first entry within 0.1 ms, else -1"""
    for i, (ti, _) in enumerate(table):
        if abs(t - ti) < 0.0001:
            return i
    return -1


def read_detected_keypoints(path: str) -> np.ndarray:
    """The parsing half of `Frame::LoadDetectedKeypointFromFile` (src/frame.cpp:222-240): comma-separated fields per line,
    the point is (field 1, field 2) as float; field 0 is the index."""
    pts = []
    with open(path, "r") as f:
        for line in f:
            line = line.strip("\r\n")
            if line == "":
                continue
            p = line.split(",")
            if len(p) < 3:
                raise ValueError(f"{path}: expected 'index, x, y', got {line!r}")
            pts.append((float(p[1]), float(p[2])))
    return np.asarray(pts, np.float64).astype(np.float32).reshape(-1, 2)


def filter_new_keypoints(pts: np.ndarray, mask: Optional[np.ndarray], n_new: int) -> np.ndarray:
    """The selection half of `LoadDetectedKeypointFromFile` (src/frame.cpp:247-262): in file order, skip points whose mask
    pixel `mask[int(y), int(x)]` is 0, stop after n_new."""
    if n_new <= 0 or pts.shape[0] == 0:
        return np.zeros((0, 2), np.float32)
    if mask is None:
        return pts[:n_new].copy()
    ok = mask[pts[:, 1].astype(np.int64), pts[:, 0].astype(np.int64)] != 0
    return pts[ok][:n_new].copy()


# ------------------------------------------------------------------------------------------------
# the sequence and the frame loop
# ------------------------------------------------------------------------------------------------
def read_gray(path: str) -> np.ndarray:
    """`cv::imread(path)` then the Frame constructor's conversion (src/frame.cpp:81-88): a 3-channel image goes through
    CV_RGB2GRAY although imread delivers BGR; that quirk is kept."""
    import cv2
    im = cv2.imread(path)
    if im is None:
        raise FileNotFoundError(path)
    if im.ndim == 3 and im.shape[2] == 3:
        return cv2.cvtColor(im, cv2.COLOR_RGB2GRAY)
    if im.ndim == 3 and im.shape[2] == 4:
        return cv2.cvtColor(im, cv2.COLOR_RGBA2GRAY)
    return np.ascontiguousarray(im)


@dataclass
class SequenceFrame:
    t: float
    path: str
    imu_t: np.ndarray    # samples since the previous frame (empty for the first frame)
    imu_w: np.ndarray


class RecordedSequence:
    """image_file_list.txt + imu.txt of one dataset directory, frames paired with their IMU windows in file order."""

    def __init__(self, dataset_dir: str, delay: float = 0.0):
        self.dataset_dir = dataset_dir
        self.images = read_image_file_list(dataset_dir)
        self.imu = read_imu_txt(os.path.join(dataset_dir, "imu.txt"))
        self.delay = delay

    def __len__(self):
        return len(self.images)

    def frames(self) -> Iterator[SequenceFrame]:
        feed = ImuFeed(self.imu)
        time_prev = 0.0
        for t, path in self.images:
            idx = feed.window(time_prev, t, self.delay)
            yield SequenceFrame(t, path, self.imu.t[idx], self.imu.w[idx])
            time_prev = t


@dataclass
class FrameResult:
    t: float
    n_ref: int                      # reference keypoints that went into the tracker (0 for the first frame)
    n_predict: int                  # TrackFeatures' return value
    n_carried: int                  # survivors carried over by SetPredictKeyPointsAndMask
    n_new: int                      # keypoints added by the top-up
    n_imu: int                      # IMU samples of the pair (mvImuFromLastFrame.size())
    outputs: Optional[capi.PairOutputs]
    keys_un: np.ndarray             # the frame's keypoints after carry-over and top-up: the next pair's reference keypoints
    index_in_last: np.ndarray       # for the carried ones, their index in the previous frame; -1 for new ones


def replay(ctx, seq: RecordedSequence, settings: Settings, params: Optional[capi.PagkParams] = None,
           keypoint_dir: Optional[str] = None, max_frames: Optional[int] = None, images: Optional[Sequence[np.ndarray]] = None
           ) -> Iterator[FrameResult]:
    """The demo's frame loop (Examples/Demo/RealSenseD435i.cpp:198-296) on the device: for every frame after the first
    `TrackFeatures` (stream continuation: only the new image crosses PCIe, its pyramid is the only one built) ->
    `SetPredictKeyPointsAndMask` on the resident results -> top-up to `KeyPointNumber` under the occupancy mask.

    The top-up is per-cell FAST (`ORBextractor::DetectFeatures`' detection half, strongest responses first; the reference's
    octree thinning is not reproducible) or, with `keypoint_dir`, the SuperPoint files through
    corresponds.txt as `LoadDetectedKeypoints: 1` does.  `GeometryValidation` needs the two RANSAC models from the caller and is
    left out here (`Context.geometry_validation` takes them).  New keypoints are taken as undistorted coordinates, as
    `LoadDetectedKeypointFromFile` takes the SuperPoint points (exact for rectified input, which is what both drivers feed
    the tracker).  `images` overrides the decoded files (tests).  `ctx` is a `tracker.Context`; anything with its
    `track_batch`, `set_predict_keypoints_and_mask` and `orb_cell_detect` works (the tests drive the loop with the oracle)."""
    prm = params or capi.default_params(pyramids=3, half_patch=settings.half_patch_size)
    K, Rbc, dist = settings.K, settings.Rbc, settings.dist
    n_want = settings.keypoint_number
    table = read_time_correspondences(os.path.join(keypoint_dir, "corresponds.txt")) if keypoint_dir else None
    inv = np.array([np.float32(1.0) / K[0, 0], np.float32(1.0) / K[1, 1]], np.float32)
    c0 = np.array([K[0, 2], K[1, 2]], np.float32)

    def top_up(gray, t, mask, n_have):
        n_new = n_want - n_have
        if n_new <= 0:
            return np.zeros((0, 2), np.float32)
        if table is not None:
            i = find_time_correspondence_index(table, t)
            if i < 0:
                return np.zeros((0, 2), np.float32)
            return filter_new_keypoints(read_detected_keypoints(os.path.join(keypoint_dir, table[i][1] + ".txt")), mask, n_new)
        xy, rs = ctx.orb_cell_detect(gray, 20, 7, mask=mask)
        order = np.argsort(-rs, kind="stable")[:n_new]
        return xy[order].astype(np.float32)

    keys = np.zeros((0, 2), np.float32)
    prev_t, first = 0.0, True
    for k, fr in enumerate(seq.frames()):
        if max_frames is not None and k >= max_frames:
            break
        gray = images[k] if images is not None else read_gray(fr.path)
        if first:
            new = top_up(gray, fr.t, None, 0)
            keys = new
            yield FrameResult(fr.t, 0, 0, 0, new.shape[0], 0, None, keys.copy(), np.full(keys.shape[0], -1, np.int32))
            prev_gray, prev_t, first = gray, fr.t, False
            continue
        pair = capi.PairInputs(prev_gray if k == 1 else None, gray, keys, fr.imu_t, fr.imu_w, prev_t, fr.t, K, Rbc,
                               dist=dist, n_dist=int(dist.size))
        out = ctx.track_batch([pair], prm)[0]
        last_normal = ((keys - c0) * inv).astype(np.float32)
        c = capi.CarryCase(None, None, None, last_normal, K, fr.t, prev_t, gray.shape[1], gray.shape[0], n_keys=pair.n_keys)
        (n,) = ctx.set_predict_keypoints_and_mask([c])
        new = top_up(gray, fr.t, c.mask, n)
        keys = np.concatenate([c.keys_un[:n], new]).astype(np.float32)
        idx = np.concatenate([c.index_in_last[:n], np.full(new.shape[0], -1, np.int32)]).astype(np.int32)
        yield FrameResult(fr.t, pair.n_keys, int(out.n_predict), int(n), int(new.shape[0]), int(fr.imu_t.size), out, keys.copy(), idx)
        prev_gray, prev_t = gray, fr.t

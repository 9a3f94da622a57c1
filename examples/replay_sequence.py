#!/usr/bin/env python
"""Replay a recorded camera + IMU sequence through the CUDA tracker: the counterpart of the reference's demo
(Examples/Demo/RealSenseD435i.cpp, `./RealSenseD435i RealSenseD435i.yaml`) without its windows.

    python examples/replay_sequence.py <settings.yaml> [--dataset-dir DIR] [--keypoints DIR] [--frames N]

The settings file is the reference's own (Examples/Demo/RealSenseD435i.yaml works as it stands); `datasetDir` is taken from
it unless --dataset-dir is given; with `LoadDetectedKeypoints: 1` (or --keypoints) new keypoints come from the SuperPoint
files through corresponds.txt, otherwise from the per-cell FAST detector on the device.  Needs a CUDA device: there is no
CPU path."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, sequence, tracker  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("settings")
    ap.add_argument("--dataset-dir")
    ap.add_argument("--keypoints", help="directory with corresponds.txt and the <stem>.txt keypoint files")
    ap.add_argument("--frames", type=int)
    ap.add_argument("--pyramids", type=int, default=3)      # the tracker hard-codes 3 (src/gyro_aided_tracker.cpp:276-277)
    a = ap.parse_args()
    base = os.path.dirname(os.path.abspath(a.settings))
    s = sequence.load_configure_file(a.settings)
    d = a.dataset_dir or os.path.join(base, s.dataset_dir)
    kd = a.keypoints or (os.path.join(base, s.detected_keypoints_file.lstrip("/")) if s.load_detected_keypoints else None)
    seq = sequence.RecordedSequence(d)
    prm = capi.default_params(pyramids=a.pyramids, half_patch=s.half_patch_size)
    print(f"{len(seq)} frames, {seq.imu.t.size} IMU samples, {s.width}x{s.height}, {s.keypoint_number} keypoints, "
          f"half patch {s.half_patch_size}, new keypoints from {'files' if kd else 'per-cell FAST'}")
    with tracker.Context(max_width=s.width, max_height=s.height, max_keys=max(s.keypoint_number, 16), max_pairs=1,
                         max_levels=a.pyramids, max_half_patch=s.half_patch_size, max_imu=256) as ctx:
        t0, n_feat = time.perf_counter(), 0
        for k, r in enumerate(sequence.replay(ctx, seq, s, prm, keypoint_dir=kd, max_frames=a.frames)):
            if r.outputs is None:
                print(f"T: {r.t:.6f}, first frame, {r.n_new} keypoints")
                continue
            n_feat += r.n_ref
            # the line the reference writes to trackFeatures.txt (src/gyro_aided_tracker.cpp:485-495), minus GeometryValidation
            print(f"T: {r.t:.6f}, RefKey Num: {r.n_ref}, patchMatchPredict Num: {r.n_predict}, carried: {r.n_carried}, "
                  f"feature track rate: {100.0 * r.n_carried / max(r.n_ref, 1):.1f}%, IMU num: {r.n_imu}, "
                  f"new: {r.n_new}")
        dt = time.perf_counter() - t0
    print(f"{n_feat} features over {k} pairs in {dt:.2f} s (decoding included)")


if __name__ == "__main__":
    main()

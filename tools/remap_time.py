"""Config B's 64 pairs through the pipeline with rectification maps set (run under `ncu --metrics gpu__time_duration.sum
-k regex:remap_slots` for the device time of pagk_remap_slots_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
pairs = [synth.make_pair(2000 + i, **cfg) for i in range(64)]
prm = capi.default_params(pyramids=4)
H, W = pairs[0].img_cur.shape
ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
mx, my = (xs + 1.5 * np.sin(ys / 37.0)).astype(np.float32), (ys + 1.2 * np.cos(xs / 41.0)).astype(np.float32)
with tracker.Context(max_keys=1024, max_pairs=64, max_levels=4) as ctx:
    ctx.set_rectify_maps(mx, my)
    for _ in range(4):
        out = ctx.track_batch(pairs, prm)
    print("tracked", sum(o.n_predict for o in out))

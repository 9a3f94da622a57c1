"""debug aid: one of the test_batch_shapes cases with the mismatching features printed"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
from oracle import oracle
from tests import helpers
n_pairs, n_keys = int(sys.argv[1]), int(sys.argv[2])
pairs = [synth.make_pair(8100 + i, width=320, height=240, n_keys=n_keys, pyramids=3, border=16) for i in range(n_pairs)]
prm = capi.default_params(pyramids=3)
with tracker.Context(max_width=1920, max_height=1080, max_keys=8192, max_pairs=8, max_levels=5, max_half_patch=10, max_imu=64) as ctx:
    for rep in range(3):
        gpu = ctx.track_batch(pairs, prm)
        rc, cpu = oracle.track_batch(pairs, prm, 8)
        for k, (g, c) in enumerate(zip(gpu, cpu)):
            bad = np.nonzero(~helpers.bits_equal(g.pm_pt_un, c.pm_pt_un).all(axis=1) | (g.iters != c.iters) | ~helpers.bits_equal(g.pixel_error, c.pixel_error))[0]
            if len(bad):
                print("rep", rep, "pair", k, "bad features", len(bad))
                for i in bad[:12]:
                    print("  i", i, "key", pairs[k].keys_ref_un[i], "gpu", g.pm_pt_un[i], g.iters[i], g.pixel_error[i], g.pm_status[i],
                          "cpu", c.pm_pt_un[i], c.iters[i], c.pixel_error[i], c.pm_status[i], "affine", c.affine[i])
print("done")

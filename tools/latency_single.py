"""Latency of ONE frame pair through the blocking call (pagk_submit_batch + pagk_wait_batch on one handle, pinned host
buffers), config A: what a caller that tracks frame by frame sees.  Wall clock around the ctypes calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
name = sys.argv[1] if len(sys.argv) > 1 else "A"
cfg = {k: v for k, v in synth.CONFIGS[name].items() if k != "pairs"}
pairs = [synth.make_pair(3000 + i, **cfg) for i in range(4)]
prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=cfg["half_patch"]) if "half_patch" in cfg else capi.default_params(pyramids=cfg["pyramids"])
with tracker.Context(max_width=cfg["width"], max_height=cfg["height"], max_keys=cfg["n_keys"], max_pairs=1, max_levels=cfg["pyramids"],
                     max_half_patch=cfg.get("half_patch", 5)) as ctx:
    prepared = []
    for p in pairs:
        outs = [capi.PairOutputs(p.n_keys)]
        prepared.append((capi.make_in_array([p]), capi.make_out_array(outs), outs))
    for i in range(20):
        ins, oarr, outs = prepared[i % 4]
        ctx.submit_prepared(prm, ins, oarr, 1); ctx.wait()
    ts = []
    for i in range(300):
        ins, oarr, outs = prepared[i % 4]
        t0 = time.perf_counter()
        ctx.submit_prepared(prm, ins, oarr, 1); ctx.wait()
        ts.append(time.perf_counter() - t0)
    capi.sync_out_array(oarr, outs)
    ts = np.array(ts) * 1e6
    print(f"config {name}: one pair per blocking call (pageable host buffers, ctypes): median {np.median(ts):.1f} us, p10 {np.percentile(ts, 10):.1f}, "
          f"p90 {np.percentile(ts, 90):.1f} ({outs[0].n_predict} of {pairs[0].n_keys} tracked)")

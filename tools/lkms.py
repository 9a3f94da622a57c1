"""LK kernel milliseconds for config B (one handle, clean CUDA events); usage: lkms.py [lib.so ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
pairs = [synth.make_pair(2000 + i, **cfg) for i in range(64)]
prm = capi.default_params(pyramids=4)
libs = sys.argv[1:] or [capi.lib_path()]
for lib in libs:
    capi._lib = capi.load(lib)
    with tracker.Context(max_keys=1024, max_pairs=64, max_levels=4) as ctx:
        ctx.upload(pairs, prm)
        ms = []
        for _ in range(8):
            ctx.run(); ctx.synchronize(); ms.append(ctx.last_run_ms()["lk"])
        outs = [capi.PairOutputs(p.n_keys) for p in pairs]
        ctx.download(outs)
        it = sum(o.n_iterations for o in outs)
        print(f"{os.path.basename(lib):28s} lk ms min {min(ms[2:]):.3f} med {np.median(ms[2:]):.3f}  -> {it / min(ms[2:]) / 1e3:.1f} M feature-iter/s (kernel only)")

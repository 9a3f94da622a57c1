"""Run one resident batch of config B with PAGK_LK_TIMELINE set and summarise CTA 0's stage timeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs("gpurun_out", exist_ok=True)
os.environ["PAGK_LK_TIMELINE"] = "gpurun_out/timeline.txt"
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
cfg = {k: v for k, v in synth.CONFIGS["B"].items() if k != "pairs"}
pairs = [synth.make_pair(2000 + i, **cfg) for i in range(64)]
prm = capi.default_params(pyramids=4)
with tracker.Context(max_keys=1024, max_pairs=64, max_levels=4) as ctx:
    ctx.upload(pairs, prm)
    for _ in range(3):
        ctx.run(); ctx.synchronize()
        print("run ms", ctx.last_run_ms())
hdr = open("gpurun_out/timeline.txt").readline().split()[1:]
cy = np.array([int(x.split(":")[0]) for x in hdr]); st = np.array([int(x.split(":")[1]) for x in hdr])
print("per-CTA cycles: min %d mean %d max %d ; stages: min %d mean %.0f max %d ; cycles/stage mean %.0f" % (cy.min(), cy.mean(), cy.max(), st.min(), st.mean(), st.max(), (cy / st).mean()))
t = np.loadtxt("gpurun_out/timeline.txt")
t = t[5:-5]
names = ["start", "w0 acc end", "w0 bar end", "w0 solve end", "w0 consume end", "w0 A end", "w1 acc end", "w1 produce end",
         "w1 A end", "w2 A end", "w15 A end", "after barrier", "solve: llt start", "solve: llt end", "solve: update end", "-"]
rel = t - t[:, :1]
print("stages", len(t), "mean stage cycles", np.mean(t[:, 11] - t[:, 0]))
for i, n in enumerate(names):
    print(f"{n:18s} mean {rel[:, i].mean():9.0f}  p10 {np.percentile(rel[:, i], 10):9.0f}  p90 {np.percentile(rel[:, i], 90):9.0f}")

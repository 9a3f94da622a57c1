"""Phase profile of the lane-per-feature LK kernel (needs a build with PAGK_NVCC_EXTRA=-DPAGK_LANES_PROF)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs("gpurun_out", exist_ok=True)
os.environ["PAGK_LK_TIMELINE"] = "gpurun_out/lanes_prof.txt"
import numpy as np
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker
if len(sys.argv) > 1:
    capi._lib = capi.load(sys.argv[1])
CFG = os.environ.get("PROF_CONFIG", "B")  # PROF_CONFIG=A PROF_PAIRS=1: the sparse rounds of a single frame pair
cfg = {k: v for k, v in synth.CONFIGS[CFG].items() if k != "pairs"}
NP = int(os.environ.get("PROF_PAIRS", "64"))
pairs = [synth.make_pair(2000 + i, **cfg) for i in range(min(NP, 64))]
prm = capi.default_params(pyramids=cfg["pyramids"])
pairs = (pairs * ((NP + 63) // 64))[:NP]
with tracker.Context(max_width=cfg["width"], max_height=cfg["height"], max_keys=cfg["n_keys"], max_pairs=NP, max_levels=cfg["pyramids"]) as ctx:
    ctx.upload(pairs, prm)
    for _ in range(3):
        ctx.run(); ctx.synchronize()
        print("run ms", ctx.last_run_ms())
    outs = [capi.PairOutputs(p.n_keys) for p in pairs]
    ctx.download(outs)
it = np.concatenate([o.iters for o in outs])
print("features", it.size, "iterations: mean %.2f p50 %d p90 %d p99 %d max %d" % (it.mean(), np.percentile(it, 50), np.percentile(it, 90), np.percentile(it, 99), it.max()))
t = np.loadtxt("gpurun_out/lanes_prof.txt")
names = ["refill", "setup", "waiting-lane-rounds", "pass", "coop", "solve", "rounds", "lane-rounds"]
tot = t[:, [0, 1, 3, 4, 5]].sum(axis=1)
print("warps", len(t), "cycles/warp: min %d mean %d max %d" % (tot.min(), tot.mean(), tot.max()))
for i, n in enumerate(names):
    print(f"{n:12s} mean {t[:, i].mean():10.0f}  min {t[:, i].min():10.0f} max {t[:, i].max():10.0f}" + (f"  share {100 * t[:, i].sum() / tot.sum():5.1f}%" if i in (0, 1, 3, 4, 5) else ""))
print("slots per warp-round: active %.2f waiting %.2f" % (t[:, 7].sum() / t[:, 6].sum(), t[:, 2].sum() / t[:, 6].sum()))
print("lane occupancy %.3f ; cycles per round %.0f ; pass cycles per round %.0f" % (t[:, 7].sum() / (32 * t[:, 6].sum()), tot.sum() / t[:, 6].sum(), t[:, 3].sum() / t[:, 6].sum()))

if t.shape[1] >= 16:
    t0 = t[:, 8].min()
    st, ex, en = (t[:, 8] - t0) / 1e3, (t[:, 9] - t0) / 1e3, (t[:, 10] - t0) / 1e3
    print("wall clock (us from the first warp's start): start max %.1f | queue exhausted min %.1f mean %.1f max %.1f | end min %.1f mean %.1f max %.1f" %
          (st.max(), ex.min(), ex.mean(), ex.max(), en.min(), en.mean(), en.max()))
    rb, lb = t[:, 11], t[:, 12]
    ra, la = t[:, 6] - rb, t[:, 7] - lb
    print("before the queue ran out: rounds/warp %.1f, lane occupancy %.3f ; after: rounds/warp %.1f, lane occupancy %.3f" %
          (rb.mean(), lb.sum() / (32 * rb.sum()), ra.mean(), la.sum() / (32 * max(ra.sum(), 1))))
    print("coop slot-passes/warp %.1f (%.1f %% of lane-rounds) ; sparse rounds/warp %.1f ; window stagings/warp %.1f" %
          (t[:, 13].mean(), 100 * t[:, 13].sum() / t[:, 7].sum(), t[:, 14].mean(), t[:, 15].mean()))
    print("end-time percentiles (us):", " ".join("%.0f" % np.percentile(en, q) for q in (1, 10, 25, 50, 75, 90, 99, 100)))

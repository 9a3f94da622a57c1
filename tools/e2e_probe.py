"""Where does the end-to-end step go?  host time of submit / wait, and the raw H2D rate of the same pinned block."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, tracker
ND = 5
batches, cfg = bench.make_batches("B", 64, ND, 2000, True)
N = cfg["n_keys"]; prm = capi.default_params(pyramids=4)
ctx = [tracker.Context(max_keys=N, max_pairs=64, max_levels=4) for _ in range(ND)]
obl = [bench.OutBlock(64, N) for _ in range(ND)]
ins = [capi.make_in_array(b["pairs"]) for b in batches]; oarr = [capi.make_out_array(o.outs) for o in obl]
# raw copy rate
t = torch.from_numpy(batches[0]["imgs"]); d = torch.empty_like(t, device="cuda")
for _ in range(3): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("raw H2D of the image block: %.3f ms = %.1f GB/s" % (dt * 1e3, t.numel() / dt / 1e9))
def loop(steps, depth):
    sub, wai = [], []
    infl = [False] * ND
    t0 = time.perf_counter()
    for k in range(steps):
        j = k % depth
        if infl[j]:
            a = time.perf_counter(); ctx[j].wait(); wai.append(time.perf_counter() - a)
        a = time.perf_counter(); ctx[j].submit_prepared(prm, ins[j], oarr[j], 64); sub.append(time.perf_counter() - a)
        infl[j] = True
    for j in range(depth):
        if infl[j]: ctx[j].wait()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dt, np.mean(sub), np.mean(wai) if wai else 0
for depth in (1, 2, 3, 4, 5):
    loop(6, depth)
    dt, s, w = loop(40, depth)
    print("depth %d: %.3f ms/step (%.1f M feat/s)  submit host %.3f ms  wait %.3f ms" % (depth, dt * 1e3, 64 * N / dt / 1e6, s * 1e3, w * 1e3))

#!/usr/bin/env python
"""bench.py -- throughput of the pixel-aware gyro-aided KLT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU reference arm (oracle port)

A "step" is one pass of the hot path (pyramids, gyro prediction, coarse-to-fine patch alignment,
filter) over one batch of BASELINE.json config B: 64 frame pairs of 752x480 with 1024 features
each, 4 pyramid levels, 11x11 patches, eType 4 (illumination + affine deformation).  One JSON line
is printed by rank 0.  Under torchrun every rank owns one GPU and its own batches (independent
streams, no collective on the data path: "weak" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth  # noqa: E402

METRIC = "tracked features/sec at 752x480, 1024 feats, 4 levels; px error vs CPU ref"
# algorithmic work of one feature-iteration at an 11x11 patch
FP32_FLOP_PER_FEATURE_ITER = {5: 8.5e3, 10: 30.9e3}
HBM_BYTES_PER_FEATURE_ITER = 120.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the patch-alignment kernels on a workload: NOT measured in this
# run; read from the committed summary of the `ncu --set full` capture (profiles/r02_k3_traffic.json), null when absent
def profiled_traffic(cfg_name, n_pairs):
    p = os.path.join(ROOT, "profiles", "r02_k3_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        e = d.get(f"{cfg_name}/{n_pairs}")
        return (e["bytes_per_launch"], "profiles/r02_k3_traffic.json: " + e["note"]) if e else (None, None)
    except Exception:
        return None, None
DEVICE_SHARE = int(os.environ.get("PAGK_BENCH_SHARE", "3"))  # pagk_set_device_share of the handles of the resident pipelined leg
# ... and of the end-to-end legs, whose handles wait for their downloads: fewer launches are in flight at once (measured:
# 98-101 M features/s over camera streams with 2, 95.5 M with 3)
E2E_SHARE = int(os.environ.get("PAGK_BENCH_E2E_SHARE", "2"))
N_ROTATE = int(os.environ.get("PAGK_BENCH_ROTATE", "4"))  # resident batches per GPU; 4 x 61 MB of pyramids > 126 MB L2
E2E_DEPTH = 5  # handles (streams) the end-to-end leg rotates over: uploads, kernels and downloads of 5 batches in flight


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """Everything any library prints on fd 1 (NCCL's version banner, ...) goes to stderr; the JSON line alone is
    written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """polls NVML (SM clock, throttle reasons) while the timed region runs"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pinned(shape, dtype):
    """page-locked host array (torch is only the allocator here)"""
    import torch
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=torch.cuda.is_available())
    return t.numpy()


def make_batches(cfg_name: str, n_pairs: int, n_batches: int, seed0: int, use_pinned: bool):
    """n_batches host batches; images/keypoints of a batch live in one contiguous (pinned) block"""
    from concurrent.futures import ThreadPoolExecutor
    cfg = {k: v for k, v in synth.CONFIGS[cfg_name].items() if k != "pairs"}
    H, W, N = cfg["height"], cfg["width"], cfg["n_keys"]
    alloc = pinned if use_pinned else (lambda s, d: np.empty(s, d))
    batches = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for b in range(n_batches):
            imgs = alloc((n_pairs, 2, H, W), np.uint8)
            keys = alloc((n_pairs, N, 2), np.float32)
            raw = list(ex.map(lambda i: synth.make_pair(seed0 + b * n_pairs + i, **cfg), range(n_pairs)))
            pairs = []
            for i, p in enumerate(raw):
                imgs[i, 0], imgs[i, 1], keys[i] = p.img_ref, p.img_cur, p.keys_ref_un
                q = capi.PairInputs(imgs[i, 0], imgs[i, 1], keys[i], p.imu_t, p.imu_w, p.t_ref, p.t_cur, p.K, p.Rbc,
                                    dist=p.dist, n_dist=p.n_dist)
                pairs.append(q)
            batches.append(dict(pairs=pairs, imgs=imgs, keys=keys))
    return batches, cfg


class OutBlock:
    """result vectors of a batch in contiguous pinned blocks (what SetBackToFrame copies out,
    reference src/gyro_aided_tracker.cpp:97-111, plus the patch-match status/error)"""
    FIELDS = [("pt_predict_un", np.float32, (2,)), ("pt_predict", np.float32, (2,)), ("status", np.uint8, ()),
              ("pt_gyro_predict_un", np.float32, (2,)), ("ncc", np.float32, ()), ("corner_flows", np.float32, (4, 2))]

    def __init__(self, n_pairs, n_keys, use_pinned=True, all_fields=False):
        alloc = pinned if use_pinned else (lambda s, d: np.zeros(s, d))
        self.outs = [capi.PairOutputs.__new__(capi.PairOutputs) for _ in range(n_pairs)]
        fields = capi._OUT_SPEC if all_fields else self.FIELDS
        self.blocks = {}
        for o in self.outs:
            o.n_keys, o.struct = n_keys, capi.PagkPairOut()
        for name, dt, tail in fields:
            blk = alloc((n_pairs, n_keys) + tuple(tail), dt)
            blk[...] = 0
            self.blocks[name] = blk
            for i, o in enumerate(self.outs):
                setattr(o, name, blk[i])
                setattr(o.struct, name, capi._ptr(blk[i], capi._PTR_OF[dt]))
        self.nbytes = sum(b.nbytes for b in self.blocks.values())


def cpu_arm():
    """the CPU implementation the CPU legs time and check against: oracle/_ref/libpagk_ref.so when it was built (the
    reference's own src/gyro_aided_tracker.cpp + src/patch_match.cpp + src/utils.cpp compiled in the build container
    against stand-in OpenCV/Eigen/glog headers; the prebuilt library travels with the snapshot), else the restatement"""
    from oracle import reference
    if reference.build() is not None:
        reference.load()
        return reference, "reference", ("reference sources compiled against stand-in OpenCV/Eigen/glog headers "
                                        "(oracle/ref_shim), cv::parallel_for_ = static split over std::threads")
    from oracle import oracle
    oracle.build()
    return oracle, "port", "restatement oracle/pagk_oracle.cpp"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same config."""
    if rank != 0:
        return
    oracle, kind, kind_note = cpu_arm()
    cores = os.cpu_count() or 1
    sample_pairs = args.ref_pairs or args.pairs or min(synth.CONFIGS[args.config]["pairs"], 64)
    distinct = min(sample_pairs, 16)   # generating synthetic pairs is the slow part: the batch repeats 16 distinct pairs
    batches, cfg = make_batches(args.config, distinct, 1, 1000 * (ord(args.config) - 64), False)
    batches[0]["pairs"] = [batches[0]["pairs"][i % distinct] for i in range(sample_pairs)]
    pairs = batches[0]["pairs"]
    prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=cfg["half_patch"])
    for _ in range(max(1, args.warmup // 3)):
        oracle.track_batch(pairs[:2], prm, cores)
    t0 = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        rc, outs = oracle.track_batch(pairs, prm, cores)
        iters += sum(o.n_iterations for o in outs)
    dt = time.perf_counter() - t0
    feats = args.steps * sample_pairs * cfg["n_keys"]
    v = feats / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "features/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg, sample_pairs)},
            "feature_iterations_per_sec": iters / dt,
            "cpu_baseline": {"value": v, "unit": "features/s", "cores": cores, "kind": kind, "kind_note": kind_note,
                             "sample": f"{sample_pairs} frame pairs per step x {args.steps} steps, {cores} std::threads over features"},
            "e2e": {"value": v, "unit": "features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def stream_leg(args, cfg, prm, n_streams, new_ctx, barrier, world):
    """End to end over camera streams: every handle owns n_streams streams; a step uploads ONE new frame per stream (img_ref =
    NULL: the previous current image and its pyramid are on the device), keypoints and gyro, and reads the results back."""
    from concurrent.futures import ThreadPoolExecutor
    import torch
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import sharding
    steps = args.e2e_steps or max(3, min(args.steps, 30))
    per_handle = -(-(steps + E2E_DEPTH + 2) // E2E_DEPTH)        # continuation steps each handle plays (warm-up included)
    T = per_handle + 2
    H, W, N = cfg["height"], cfg["width"], cfg["n_keys"]
    n_distinct = min(n_streams, 8)                               # distinct synthetic streams; the others replay them
    kw = {k: v for k, v in cfg.items() if k not in ("pairs",)}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        seqs = list(ex.map(lambda i: synth.make_sequence(40000 + i, T, **kw), range(n_distinct)))
    frames = pinned((T, n_streams, H, W), np.uint8)
    keys = pinned((T - 1, n_streams, N, 2), np.float32)
    for s_ in range(n_streams):
        fr, prs = seqs[s_ % n_distinct]
        for t in range(T):
            frames[t, s_] = fr[t]
        for t in range(T - 1):
            keys[t, s_] = prs[t].keys_ref_un
    def batch(t, cont):
        prs = [seqs[s_ % n_distinct][1][t] for s_ in range(n_streams)]
        return [capi.PairInputs(None if cont else frames[t, s_], frames[t + 1, s_], keys[t, s_], q.imu_t, q.imu_w, q.t_ref, q.t_cur,
                                q.K, q.Rbc) for s_, q in enumerate(prs)]
    ctxs = [new_ctx() for _ in range(E2E_DEPTH)]
    oblocks = [OutBlock(n_streams, N) for _ in range(E2E_DEPTH)]
    oarrs = [capi.make_out_array(ob.outs) for ob in oblocks]
    for c in ctxs:
        c.set_stage_timing(False)
        c.set_device_share(E2E_SHARE)
    keep = [batch(0, False)] + [batch(t, True) for t in range(1, T - 1)]
    ins = [capi.make_in_array(b) for b in keep]
    pos = [0] * E2E_DEPTH                                        # next frame pair of each handle's streams
    inflight = [False] * E2E_DEPTH

    def run_steps(k0, k1):
        for k in range(k0, k1):
            j = k % E2E_DEPTH
            if inflight[j]:
                ctxs[j].wait()
            ctxs[j].submit_prepared(prm, ins[pos[j]], oarrs[j], n_streams)
            pos[j] += 1
            inflight[j] = True

    def drain():
        for j in range(E2E_DEPTH):
            if inflight[j]:
                ctxs[j].wait()
                inflight[j] = False
    warm = E2E_DEPTH + 2          # every handle has its first (two-image) batch behind it: all timed steps are continuations
    run_steps(0, warm)
    drain()
    barrier()
    t0 = time.perf_counter()
    run_steps(warm, warm + steps)
    drain()
    torch.cuda.synchronize()
    dte = sharding.reduce_time_max(time.perf_counter() - t0)
    tracked = int(sum(int(o.n_predict) for o in oarrs[0][:n_streams]))
    h2d = int(frames[0].nbytes + keys[0].nbytes + n_streams * 96)
    res = {"value": world * steps * n_streams * N / dte, "unit": "features/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(oblocks[0].nbytes + n_streams * 24), "steps": steps, "streams_per_step": n_streams,
           "tracked_in_last_step_of_handle_0": tracked,
           "api": f"pagk_submit_batch/pagk_wait_batch with img_ref = NULL (stream continuation) over {E2E_DEPTH} handles"}
    for c in ctxs:
        c.close()
    return res


def bind_to_gpu_numa_node(index: int):
    """run this process (its pinned allocations are first-touched by it) on the CPUs NVML names as local to the GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_threads(args):
    """--threads: one process, one host thread per GPU, every thread with its own handles on its own device (include/pagk.h:
    "one handle per device; handles on different devices are independent").  Resident and end-to-end legs, reduced line."""
    import torch
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    n_dev = args.gpus
    if torch.cuda.device_count() < n_dev:
        raise SystemExit(f"bench.py --threads: {n_dev} GPUs asked for, {torch.cuda.device_count()} visible")
    cfgd = synth.CONFIGS[args.config]
    n_pairs = args.pairs or min(cfgd["pairs"], 64)
    batches, cfg = make_batches(args.config, n_pairs, N_ROTATE, 1000 * (ord(args.config) - 64), True)
    N, half = cfg["n_keys"], cfg["half_patch"]
    prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=half)
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 30))
    gate = threading.Barrier(n_dev)
    res = [None] * n_dev

    def worker(d):
        mk = lambda: tracker.Context(device=d, max_width=cfg["width"], max_height=cfg["height"], max_keys=N, max_pairs=n_pairs,
                                     max_levels=cfg["pyramids"], max_half_patch=half)
        pctx = [mk() for _ in range(N_ROTATE)]
        for c, b in zip(pctx, batches):
            c.upload(b["pairs"], prm)
            c.set_stage_timing(False)
            c.set_device_share(DEVICE_SHARE)
        for k in range(max(3, args.warmup) + N_ROTATE):
            pctx[k % N_ROTATE].run()
        for c in pctx:
            c.synchronize()
        gate.wait()
        t0 = time.perf_counter()
        for k in range(args.steps):
            pctx[k % N_ROTATE].run()
        for c in pctx:
            c.synchronize()
        dt = time.perf_counter() - t0
        ectx = [mk() for _ in range(E2E_DEPTH)]
        for c in ectx:
            c.set_stage_timing(False)
            c.set_device_share(E2E_SHARE)
        oblocks = [OutBlock(n_pairs, N) for _ in range(E2E_DEPTH)]
        ins = [capi.make_in_array(batches[j % N_ROTATE]["pairs"]) for j in range(E2E_DEPTH)]
        oarrs = [capi.make_out_array(ob.outs) for ob in oblocks]

        def loop(steps):
            inflight = [False] * E2E_DEPTH
            for k in range(steps):
                j = k % E2E_DEPTH
                if inflight[j]:
                    ectx[j].wait()
                ectx[j].submit_prepared(prm, ins[j], oarrs[j], n_pairs)
                inflight[j] = True
            for j in range(E2E_DEPTH):
                if inflight[j]:
                    ectx[j].wait()
        loop(E2E_DEPTH + 2)
        gate.wait()
        t0 = time.perf_counter()
        loop(e2e_steps)
        dte = time.perf_counter() - t0
        res[d] = (dt, dte, int(oarrs[0][0].n_predict))
        for c in pctx + ectx:
            c.close()
    ths = [threading.Thread(target=worker, args=(d,)) for d in range(n_dev)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if any(r is None for r in res):
        raise SystemExit("bench.py --threads: a device thread failed")
    dt, dte = max(r[0] for r in res), max(r[1] for r in res)
    feats = n_pairs * N
    b0 = batches[0]
    emit({"metric": METRIC, "value": n_dev * args.steps * feats / dt, "unit": "features/s", "n_gpus": n_dev, "steps": args.steps,
          "warmup": max(3, args.warmup), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "config": {"workload": workload_name(args.config, cfg, n_pairs)},
          "mode": "one process, one host thread and one set of handles per GPU (--threads)",
          "e2e": {"value": n_dev * e2e_steps * feats / dte, "unit": "features/s",
                  "h2d_bytes_per_step": int(b0["imgs"].nbytes + b0["keys"].nbytes + n_pairs * 96), "steps": e2e_steps,
                  "tracked_in_first_pair": [r[2] for r in res]}})


def workload_name(name, cfg, n_pairs):
    return (f"config {name}: {cfg['width']}x{cfg['height']}, {cfg['n_keys']} features/pair, {cfg['pyramids']} levels, "
            f"{2 * cfg['half_patch'] + 1}x{2 * cfg['half_patch'] + 1} patch, eType 4, {n_pairs} frame pairs/step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pagk", choices=["pagk", "reference"])
    ap.add_argument("--config", default="B", choices=list(synth.CONFIGS))
    ap.add_argument("--pairs", type=int, default=None, help="frame pairs per step (default: the config's batch, 64 for B)")
    ap.add_argument("--ref-pairs", type=int, default=None,
                    help="frame pairs per step of the CPU reference arm (default: the same batch as the CUDA arm, 64 for config B)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-stream", action="store_true",
                    help="skip the leg `e2e_stream` (the end-to-end pipeline over camera streams, BASELINE config D's unit: each step "
                         "tracks frame t against frame t-1 of every stream, whose pyramid stayed on the device)")
    ap.add_argument("--threads", action="store_true",
                    help="ONE process, one host thread and one set of handles per GPU (the in-process form of include/pagk.h) "
                         "instead of one process per GPU: prints a reduced line (value, e2e, e2e_stream)")
    args = ap.parse_args()

    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.threads:
        run_threads(args)
        return

    import torch
    import torch.distributed as dist
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import sharding, tracker
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the pagk hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None   # pinned buffers on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfgd = synth.CONFIGS[args.config]
    n_pairs = args.pairs or min(cfgd["pairs"], 64)
    t_gen = time.time()
    batches, cfg = make_batches(args.config, n_pairs, N_ROTATE, 1000 * (ord(args.config) - 64) + 100000 * rank, True)
    log(f"[rank {rank}] generated {N_ROTATE} x {n_pairs} synthetic pairs in {time.time() - t_gen:.1f}s")
    N, half = cfg["n_keys"], cfg["half_patch"]
    prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=half)
    def new_ctx():
        return tracker.Context(device=local_rank, max_width=cfg["width"], max_height=cfg["height"], max_keys=N,
                               max_pairs=n_pairs, max_levels=cfg["pyramids"], max_half_patch=half)
    ctxs = [new_ctx() for _ in range(N_ROTATE)]
    for c in ctxs[1:]:
        c.share_stream(ctxs[0])      # one in-order stream: steps cannot overlap, CUDA events stay clean
    feats_per_step = n_pairs * N

    # ---- resident leg: inputs already in HBM, kernels only -------------------------------------
    for c, b in zip(ctxs, batches):
        c.upload(b["pairs"], prm)
    for k in range(max(3, args.warmup)):
        ctxs[k % N_ROTATE].run()
    ctxs[0].synchronize()
    outs0 = OutBlock(n_pairs, N, all_fields=True)
    ctxs[0].run()
    ctxs[0].download(outs0.outs)
    iters_per_step = [None] * N_ROTATE
    for k, c in enumerate(ctxs):
        ob = OutBlock(n_pairs, N)
        c.run()
        c.download(ob.outs)
        iters_per_step[k] = sum(o.n_iterations for o in ob.outs)
    # per-stage device times from one untimed run with the stage clocks on; the timed loop runs without them (a CUDA
    # event between two kernels costs about 3 us of stream time), keeping only the two events around each LK launch
    ctxs[0].run()
    stage_ms = ctxs[0].last_run_ms()
    for c in ctxs:
        c.set_stage_timing(False)
    ctxs[0].run()
    ctxs[0].synchronize()
    launches0 = sum(c.launch_count() for c in ctxs)
    for c in ctxs:
        c.timing_reset()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ctxs[k % N_ROTATE].run()
    ctxs[0].synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks_serial = sampler.result()
    # device time of the dominant kernel: CUDA events around every LK launch of the timed region
    lk_n, lk_sum, lk_iters = 0, 0.0, 0.0
    for k, c in enumerate(ctxs):
        n_r, ms_sum = c.timing_read()
        lk_n += n_r; lk_sum += ms_sum; lk_iters += n_r * iters_per_step[k]
    dt_serial = sharding.reduce_time_max(dt)
    # ---- pipelined resident leg: the same steps, every resident batch on its own handle AND stream, so the kernels of
    # consecutive steps overlap (the tail of one step's alignment kernel runs beside the head of the next step)
    pctx = [new_ctx() for _ in range(N_ROTATE)]
    for c, b in zip(pctx, batches):
        c.upload(b["pairs"], prm)
        c.set_stage_timing(False)
        c.set_device_share(DEVICE_SHARE)   # the handles of this leg keep the device busy together (see include/pagk.h)
    for k in range(max(3, args.warmup) + N_ROTATE):
        pctx[k % N_ROTATE].run()
    for c in pctx:
        c.synchronize()
    launches0 = sum(c.launch_count() for c in pctx)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for k in range(args.steps):
        pctx[k % N_ROTATE].run()
    for c in pctx:
        c.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.result()
    launches = sum(c.launch_count() for c in pctx) - launches0
    dt_max = sharding.reduce_time_max(dt)
    total_iters = sum(iters_per_step[k % N_ROTATE] for k in range(args.steps))
    (all_iters,) = sharding.reduce_counts(total_iters)
    value = world * args.steps * feats_per_step / dt_max
    fi_per_s = all_iters / dt_max

    # ---- end-to-end leg: pinned host buffers in, pinned host buffers out, through the public C-ABI -------
    # pagk_submit_batch / pagk_wait_batch on E2E_DEPTH handles with their own streams: the uploads of the next
    # batches overlap the kernels of the current one.  Every step copies its images, keypoints and gyro data host->device
    # and its result vectors device->host inside the timed region.
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 30))
    ectx = [new_ctx() for _ in range(E2E_DEPTH)]
    for c in ectx:
        c.set_stage_timing(False)   # a throughput pipeline does not read per-stage clocks (no event between kernels)
        c.set_device_share(E2E_SHARE)
    oblocks = [OutBlock(n_pairs, N) for _ in range(E2E_DEPTH)]
    ins = [capi.make_in_array(batches[j % N_ROTATE]["pairs"]) for j in range(E2E_DEPTH)]
    oarrs = [capi.make_out_array(ob.outs) for ob in oblocks]

    def e2e_loop(steps):
        inflight = [False] * E2E_DEPTH
        for k in range(steps):
            j = k % E2E_DEPTH
            if inflight[j]:
                ectx[j].wait()
            ectx[j].submit_prepared(prm, ins[j], oarrs[j], n_pairs)
            inflight[j] = True
        for j in range(E2E_DEPTH):
            if inflight[j]:
                ectx[j].wait()
    e2e_loop(E2E_DEPTH + 2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    dte = sharding.reduce_time_max(time.perf_counter() - t0)
    e2e_value = world * e2e_steps * feats_per_step / dte
    # the end-to-end results of batch 0 must be the resident leg's results of the same batch, bit for bit
    e2e_ok = int(oarrs[0][0].n_predict) > 0 and all(
        np.array_equal(oblocks[0].outs[i].status, outs0.outs[i].status) and
        np.array_equal(oblocks[0].outs[i].pt_predict_un.view(np.uint32), outs0.outs[i].pt_predict_un.view(np.uint32))
        for i in range(n_pairs))
    b0 = batches[0]
    h2d = int(b0["imgs"].nbytes + b0["keys"].nbytes + n_pairs * 96)
    d2h = int(oblocks[0].nbytes + n_pairs * 24)

    e2e_stream = None if args.no_stream else stream_leg(args, cfg, prm, n_pairs, new_ctx, barrier, world)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the LK patch-alignment kernel) ------------------------------
    peaks, peaks_kind = measured_peaks()
    lk_avg_ms = lk_sum / max(1, lk_n)
    it_launch = lk_iters / max(1, lk_n)
    fp32_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12  # TFLOP/s with FMA, nominal lanes x clock
    ach = it_launch * FP32_FLOP_PER_FEATURE_ITER.get(half, 70.0 * (2 * half + 1) ** 2) / (lk_avg_ms * 1e-3) / 1e12
    hbm_ach = it_launch * HBM_BYTES_PER_FEATURE_ITER / (lk_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "fp32", "kernel": "pagk_lk_template_kernel + pagk_lk_lanes_kernel (K3a + K3b, one event pair around both)", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak, "frac_of_non_fma_peak": ach / (fp32_peak / 2),
                "traffic": profiled_traffic(args.config, n_pairs)[0], "traffic_from_profile": profiled_traffic(args.config, n_pairs)[1],
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum; not measured in this run)",
                "algorithmic_flop_per_feature_iteration": FP32_FLOP_PER_FEATURE_ITER.get(half, 70.0 * (2 * half + 1) ** 2),
                "peak_source": f"148 SM x 128 lanes x 2 x {peaks.get('sm_max_mhz', 1965.0):.0f} MHz ({peaks_kind} sm_max_mhz); "
                               "FMA contraction is forbidden by bit-parity, so half of it is the reachable ceiling",
                "kernel_ms": lk_avg_ms, "kernel_launches_timed": lk_n, "feature_iterations_per_launch": it_launch,
                "kernel_ms_leg": "serial: one launch at a time with the device to itself (compare with serial.ms_per_step, not with "
                                 "ms_per_step: in the pipelined leg the launches of different handles run beside each other, so a step "
                                 "takes less wall time than one launch lasts)",
                "kernel_feature_iterations_per_sec": it_launch / (lk_avg_ms * 1e-3),
                "hbm": {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                        "peak_source": peaks_kind}}
    # the same accounting over the pipelined leg as a whole: launches of several handles run beside each other there, so a
    # launch has no duration of its own -- algorithmic flops of the timed region / its wall time, per GPU, K1, K2, K3a and K4
    # included in the time (a lower bound of what the alignment kernels reach in that configuration)
    pipe_ach = (fi_per_s / world) * FP32_FLOP_PER_FEATURE_ITER.get(half, 70.0 * (2 * half + 1) ** 2) / 1e12
    roofline["pipelined_whole_step"] = {"achieved": pipe_ach, "unit": "TFLOP/s", "frac": pipe_ach / fp32_peak,
                                        "frac_of_non_fma_peak": pipe_ach / (fp32_peak / 2),
                                        "note": f"feature-iterations of the timed region of `value` x flop per unit / wall time, per GPU; "
                                                f"{N_ROTATE} handles with pagk_set_device_share({DEVICE_SHARE})"}

    # ---- CPU baseline + parity on the first batch (rank 0, bounded sample) ------------------------------
    cpu_baseline, parity = None, None
    if not args.no_cpu:
        from tests import helpers
        oracle, kind, kind_note = cpu_arm()
        cores = os.cpu_count() or 1
        sample = batches[0]["pairs"][:min(n_pairs, 32)]
        oracle.track_batch(sample[:2], prm, cores)
        t0 = time.perf_counter()
        rc, cpu = oracle.track_batch(sample, prm, cores)
        dtc = time.perf_counter() - t0
        cpu_iters = sum(o.n_iterations for o in cpu)
        # one thread on the first pairs of the same sample, and the three spans the reference times itself
        # (mTimeCostGyroPredict, mTimeCostOptFlow, mTimeCostOptFlowResultFilterOut), mean per frame pair of the threaded run
        one = sample[:min(len(sample), 4)]
        t0 = time.perf_counter()
        oracle.track_batch(one, prm, 1)
        dt1 = time.perf_counter() - t0
        spans = {k: 1e3 * float(np.mean([getattr(o.struct, k) for o in cpu])) for k in ("t_gyro_predict", "t_opt_flow", "t_filter")}
        cpu_baseline = {"value": len(sample) * N / dtc, "unit": "features/s", "cores": cores, "kind": kind,
                        "kind_note": kind_note, "feature_iterations_per_sec": cpu_iters / dtc,
                        "sample": f"first {len(sample)} frame pairs of the timed workload, {cores} std::threads over features",
                        "one_thread": {"value": len(one) * N / dt1, "unit": "features/s", "sample": f"first {len(one)} frame pairs"},
                        "reference_spans_ms_per_pair": spans}
        worst, same, tot, bit = 0.0, 0, 0, True
        for g, c in zip(outs0.outs, cpu):
            rep = helpers.compare(g, c)
            if kind == "reference":
                rep.pop("iters")   # per-feature pass counts are not observable from outside the reference; the total is
            bit &= all(v.get("bit_mismatch", 0) == 0 for v in rep.values() if isinstance(v, dict) and "bit_mismatch" in v)
            bit &= rep["Rcl_bits"] == 0 and rep["KRKinv_bits"] == 0
            ok = (c.status == 1) & (g.status == 1)
            if ok.any():
                worst = max(worst, float(np.hypot(*(g.pt_predict_un[ok] - c.pt_predict_un[ok]).T).max()))
            same += int((g.status == c.status).sum()); tot += c.status.size
        parity = {"max_px_error_vs_cpu_ref": worst, "status_equal_frac": same / max(1, tot), "bit_exact_all_outputs": bool(bit),
                  "pairs_checked": len(sample), "iterations_equal": bool(sum(g.n_iterations for g in outs0.outs[:len(sample)]) == cpu_iters)}

    line = {"metric": METRIC, "value": value, "unit": "features/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg, n_pairs)},
            "method": {"l2": f"{N_ROTATE} rotating resident batches per GPU ({(N_ROTATE * 2 * n_pairs * cfg['width'] * cfg['height'] * 4 // 3) >> 20} MiB of pyramids) > 126 MB L2",
                       "timing": f"wall clock around K back-to-back steps over {N_ROTATE} resident batches, each on its own handle and stream with pagk_set_device_share({DEVICE_SHARE}) (consecutive steps run beside each other on the device; the end-to-end legs with share {E2E_SHARE}), barrier + device synchronize on both sides, max over ranks; `serial` repeats the K steps on one in-order stream with the device to each launch (share 1), and the kernel ms of `roofline` are CUDA events around every LK launch of that serial leg; stage_ms from one untimed run with per-stage events on"},
            "feature_iterations_per_sec": fi_per_s,
            "serial": {"ms_per_step": 1e3 * dt_serial / args.steps, "value": world * args.steps * feats_per_step / dt_serial,
                       "unit": "features/s", "clocks": clocks_serial,
                       "note": "the same steps on ONE in-order stream (no overlap between steps): the leg the kernel events of `roofline` and `stage_ms` come from"},
            "stage_ms": stage_ms, "clocks": clocks, "host_cpus_bound_to_gpu_numa_node": numa_cpus,
            "e2e": {"value": e2e_value, "unit": "features/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "results_ok": bool(e2e_ok), "d2h_fields": [f[0] for f in OutBlock.FIELDS] + ["n_predict", "n_iterations", "Rcl"],
                    "api": f"pagk_submit_batch/pagk_wait_batch over {E2E_DEPTH} handles (pinned host buffers in and out, pagk_set_stage_timing off)"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
    if e2e_stream is not None:
        line["e2e_stream"] = e2e_stream
    emit(line)
    for c in (ctxs + pctx + ectx)[::-1]:   # borrowers of a shared stream before its owner
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

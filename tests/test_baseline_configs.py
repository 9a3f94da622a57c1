"""Every BASELINE.json configuration at FULL size through the C-ABI, against the reference's own sources compiled in the
build container (oracle/_ref/libpagk_ref.so; the restatement where that library is absent).

  A  640x480, 500 features, 3 levels: literally GyroAidedTracker(...).TrackFeatures() (reference src/gyro_aided_tracker.cpp:344-426
     with its hard-coded 3 levels / 10 iterations, :276-278), `last_path() == 0`
  B  752x480, 1024 features, 4 levels, a 64-pair batch                       (tests/test_gpu_parity.py has more of B)
  C  1920x1080, 8192 features, 21x21 patches, affine deformation, 4 levels
  D  256 camera+IMU streams of B sharded over the visible devices with sharding.streams_of_rank, bit-identical to one device
  E  3840x2160, 32768 features, 5 levels, 3 rad/s rotations
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, sharding, synth, tracker
from tests import helpers

pytestmark = pytest.mark.gpu


def _cpu_arm():
    from oracle import reference
    if reference.available():
        reference.load()
        return reference, True
    from oracle import oracle
    oracle.build()
    oracle.load()
    return oracle, False


def _ctx(cfg, n_pairs, device=0):
    return tracker.Context(device=device, max_width=cfg["width"], max_height=cfg["height"], max_keys=cfg["n_keys"], max_pairs=n_pairs,
                           max_levels=cfg["pyramids"], max_half_patch=cfg["half_patch"])


def _compare(gpu, cpu, is_ref):
    fields = (set(helpers.FLOAT_FIELDS) | {"status", "pm_status"}) if is_ref else None  # per-feature pass counts: restatement only
    for g, c in zip(gpu, cpu):
        helpers.assert_north_star(g, c)
        helpers.assert_bit_exact(g, c, fields=fields)
        assert g.n_predict == c.n_predict and g.n_iterations == c.n_iterations
        assert helpers.bits_equal(g.Rcl, c.Rcl).all() and helpers.bits_equal(g.KRKinv, c.KRKinv).all()


@pytest.mark.parametrize("name,n_pairs", [("A", 2), ("B", 64), ("C", 1), ("E", 1)])
def test_config_at_full_size(cuda_lib, name, n_pairs):
    cpu, is_ref = _cpu_arm()
    distinct = min(n_pairs, 8)                       # 64 pairs = 8 distinct synthetic pairs x 8 (generation is the slow part)
    base, cfg = synth.make_config_pairs(name, n_pairs=distinct)
    pairs = [base[i % distinct] for i in range(n_pairs)]
    prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=cfg["half_patch"])
    with _ctx(cfg, n_pairs) as ctx:
        gpu = ctx.track_batch(pairs, prm)
    rc, ref = cpu.track_batch(base, prm, os.cpu_count() or 4)
    assert rc == 0
    if is_ref:   # 3 levels and 10 iterations are what TrackFeatures() hard-codes: config A goes through it, the others through PatchMatch
        assert cpu.last_path() == (0 if name == "A" else 1)
    _compare(gpu, [ref[i % distinct] for i in range(n_pairs)], is_ref)
    assert sum(g.n_predict for g in gpu) > 0.5 * n_pairs * cfg["n_keys"]


def test_config_D_256_streams_sharded_over_the_visible_devices(cuda_lib):
    """256 streams -> device s mod G (sharding.streams_of_rank), one handle and one host thread per device, batches of 64:
    every stream's results equal the one-device run of the same stream (and the CPU reference on the distinct ones)"""
    cpu, is_ref = _cpu_arm()
    n_streams, batch, distinct = 256, 64, 8
    base, cfg = synth.make_config_pairs("D", n_pairs=distinct)
    streams = [base[s % distinct] for s in range(n_streams)]
    prm = capi.default_params(pyramids=cfg["pyramids"], half_patch=cfg["half_patch"])
    n_dev = max(1, cuda_lib.pagk_device_count())
    results = [None] * n_streams
    errors = []

    def worker(rank):
        try:
            mine = list(sharding.streams_of_rank(n_streams, rank, n_dev))
            with _ctx(cfg, batch, device=rank) as ctx:
                for b0 in range(0, len(mine), batch):
                    ids = mine[b0:b0 + batch]
                    for s, o in zip(ids, ctx.track_batch([streams[s] for s in ids], prm)):
                        results[s] = o
        except Exception as e:  # pragma: no cover
            errors.append((rank, repr(e)))
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(n_dev)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert sorted(s for r in range(n_dev) for s in sharding.streams_of_rank(n_streams, r, n_dev)) == list(range(n_streams))
    # the one-device run: the distinct streams in one batch on device 0
    with _ctx(cfg, distinct) as ctx:
        one = ctx.track_batch(base, prm)
    for s in range(n_streams):
        helpers.assert_bit_exact(results[s], one[s % distinct])
    rc, ref = cpu.track_batch(base, prm, os.cpu_count() or 4)
    assert rc == 0
    _compare(one, ref, is_ref)


def test_21x21_patches_level_granular_items_two_thousand_features():
    """pagk_lk_lanes_kernel<10, *> with level-granular work items (PAGK_LK_SPLIT=1; config C alone is one wave of lanes and
    keeps a feature in its lane): 2048 features on a 960x540 image, child process because the switch is read once"""
    code = (
        "import os, numpy as np\n"
        "from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker\n"
        "from oracle import oracle\n"
        "from tests import helpers\n"
        "pairs = [synth.make_pair(2300 + i, width=960, height=540, n_keys=2048, half_patch=10, pyramids=4, border=48) for i in range(2)]\n"
        "prm = capi.default_params(pyramids=4, half_patch=10)\n"
        "with tracker.Context(max_width=960, max_height=540, max_keys=2048, max_pairs=2, max_levels=4, max_half_patch=10) as ctx:\n"
        "    gpu = ctx.track_batch(pairs, prm)\n"
        "oracle.build(); oracle.load()\n"
        "rc, cpu = oracle.track_batch(pairs, prm, os.cpu_count() or 4)\n"
        "assert rc == 0\n"
        "for g, c in zip(gpu, cpu):\n"
        "    helpers.assert_north_star(g, c); helpers.assert_bit_exact(g, c)\n"
        "print('ok', sum(g.n_iterations for g in gpu))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for split in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PAGK_LK_SPLIT=split), capture_output=True, text=True,
                           timeout=900, cwd=root)
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_11x11_both_launch_shapes_and_item_modes():
    """pagk_lk_lanes_kernel<5, *, 8> and <5, *, 12> (the launcher picks by the number of features; PAGK_LK_WARPS forces
    either) with feature-granular and level-granular items, on a batch small enough for every combination to be unusual:
    3 pairs x 700 features with border features, eType 4 and eType 3 (no affine matrix).  Child processes because the
    switches are read once."""
    code = (
        "import os, numpy as np\n"
        "from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, synth, tracker\n"
        "from oracle import oracle\n"
        "from tests import helpers\n"
        "pairs = [synth.make_pair(5200 + i, width=640, height=400, n_keys=700, pyramids=4, border=6) for i in range(3)]\n"
        "oracle.build(); oracle.load()\n"
        "tot = 0\n"
        "for et in (4, 3):\n"
        "    prm = capi.default_params(pyramids=4, e_type=et)\n"
        "    with tracker.Context(max_width=640, max_height=400, max_keys=700, max_pairs=3, max_levels=4) as ctx:\n"
        "        gpu = ctx.track_batch(pairs, prm)\n"
        "    rc, cpu = oracle.track_batch(pairs, prm, os.cpu_count() or 4)\n"
        "    assert rc == 0\n"
        "    for g, c in zip(gpu, cpu):\n"
        "        helpers.assert_north_star(g, c); helpers.assert_bit_exact(g, c)\n"
        "    tot += sum(g.n_iterations for g in gpu)\n"
        "print('ok', tot)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = set()
    for warps in ("8", "12"):
        for split in ("1", "0"):
            r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PAGK_LK_WARPS=warps, PAGK_LK_SPLIT=split),
                               capture_output=True, text=True, timeout=900, cwd=root)
            assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]
            seen.add(r.stdout.split()[1])
    assert len(seen) == 1, seen   # the same number of feature-iterations whatever the shape of the launch


def test_two_devices_one_process_two_host_threads(cuda_lib):
    """include/pagk.h: "one handle per device; handles on different devices are independent".  Function attributes (the
    dynamic shared memory opt-in of the alignment kernel) and __constant__ tables belong to a device: pagk_create sets them
    for the handle's own device, so a second device in the same process runs the same kernels"""
    if cuda_lib.pagk_device_count() < 2:
        pytest.skip("needs two CUDA devices in one process")
    pairs = [synth.make_pair(9100 + i, width=320, height=240, n_keys=300, pyramids=3, border=20) for i in range(4)]
    prm = capi.default_params(pyramids=3)
    outs, errors = [None, None], []

    def worker(d):
        try:
            with tracker.Context(device=d, max_width=320, max_height=240, max_keys=300, max_pairs=4, max_levels=3) as ctx:
                for _ in range(3):
                    outs[d] = ctx.track_batch(pairs, prm)
        except Exception as e:  # pragma: no cover
            errors.append((d, repr(e)))
    ths = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
    for a, b in zip(outs[0], outs[1]):
        helpers.assert_bit_exact(a, b)
    assert sum(o.n_predict for o in outs[1]) > 0

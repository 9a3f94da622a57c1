"""CPU, world_size 2 over gloo: the multi-GPU path partitions independent streams across ranks with no data-path
collective; only the timing (MAX) and the counts (SUM) are reduced.  The per-rank worker here is the oracle."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, sharding, synth
    from oracle import oracle
    n_streams = 5
    mine = sharding.streams_of_rank(n_streams, rank, world)
    pairs = [synth.make_pair(500 + s, width=160, height=120, n_keys=32, pyramids=3, border=12, margin=32,
                             K=synth.scaled_euroc_K(160)) for s in mine]
    rc, outs = oracle.track_batch(pairs, capi.default_params(pyramids=3), 1)
    feats = sum(p.n_keys for p in pairs)
    tot = sharding.reduce_counts(feats, sum(o.n_iterations for o in outs))
    tmax = sharding.reduce_time_max(0.25 + rank)
    q.put((rank, list(mine), [o.pt_predict_un.copy() for o in outs], tot, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_streams_and_reduce():
    sys.path.insert(0, ROOT)
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi, sharding, synth
    from oracle import oracle
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    streams = sorted(s for r in res for s in r[1])
    assert streams == list(range(5))                      # disjoint and complete
    assert res[0][3] == res[1][3] and res[0][3][0] == 5 * 32   # SUM of features identical on both ranks
    assert res[0][4] == res[1][4] == 1.25                 # MAX over ranks of the per-rank time
    # sharded results are bit-identical to the single-process run of the same streams
    for r in res:
        for s, pts in zip(r[1], r[2]):
            p = synth.make_pair(500 + s, width=160, height=120, n_keys=32, pyramids=3, border=12, margin=32,
                                K=synth.scaled_euroc_K(160))
            ref = oracle.track(p, capi.default_params(pyramids=3), 1)[1]
            assert np.array_equal(pts.view(np.uint32), ref.pt_predict_un.view(np.uint32))


def test_stream_partition_properties():
    sys.path.insert(0, ROOT)
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import sharding
    for n in (0, 1, 7, 256):
        for world in (1, 2, 4, 8):
            parts = [list(sharding.streams_of_rank(n, r, world)) for r in range(world)]
            assert sorted(s for p in parts for s in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1

"""World-size-2 gloo test (CPU) of the multi-GPU logic: block sharding of independent streams and
the detection-count all-reduce.  The per-rank 'device tick' is played by the oracle."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'] = str(rank)
    os.environ['WORLD_SIZE'] = str(world)
    os.environ['LOCAL_RANK'] = str(rank)
    from mycroft_precise_b200.dist import init_from_env, shard_range, DetectionCounter
    r, _, w = init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    lo, hi = shard_range(n_streams, rank, world)
    from golden.cases import make_pcm
    from oracle.gru import GruWeights
    from oracle.listener import run_streams
    wts = GruWeights.random(13, 20, seed=0, scale=0.3)
    local = torch.zeros(1, dtype=torch.int64)
    per_tick = []
    counter = DetectionCounter(local)
    pcm = np.stack([make_pcm(500 + s, 1024 * 16, 'noise') for s in range(lo, hi)]) if hi > lo else np.zeros((0, 1024 * 16), np.int16)
    _, _, fired = run_streams(wts, pcm, 1024, sensitivity=0.9, trigger_level=0)
    for k in range(16):
        local[0] = int(fired[:, k].sum()) if len(fired) else 0
        if k % 2:
            counter.all_reduce()
        else:
            counter.all_reduce_overlapped()        # CPU tensors: same result through the synchronous path
            counter.wait()
        per_tick.append(int(counter.total.item()))
    q.put((rank, lo, hi, per_tick))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from mycroft_precise_b200.dist import shard_range, owner_of
    for n, w in ((10, 2), (7, 4), (1_000_000, 8), (3, 8), (0, 2)):
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        for r, (lo, hi) in enumerate(ranges):
            for s in (lo, hi - 1):
                if lo <= s < hi:
                    assert owner_of(s, n, w) == r
    assert shard_range(1_000_000, 3, 8) == (375000, 500000)


def test_two_rank_count_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    n_streams = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 3, 3, 5)
    assert res[0][3] == res[1][3]                       # every rank sees the same global count
    # same as one process owning all streams
    from golden.cases import make_pcm
    from oracle.gru import GruWeights
    from oracle.listener import run_streams
    wts = GruWeights.random(13, 20, seed=0, scale=0.3)
    pcm = np.stack([make_pcm(500 + s, 1024 * 16, 'noise') for s in range(n_streams)])
    _, _, fired = run_streams(wts, pcm, 1024, sensitivity=0.9, trigger_level=0)
    assert res[0][3] == [int(fired[:, k].sum()) for k in range(16)]
    assert sum(res[0][3]) > 0

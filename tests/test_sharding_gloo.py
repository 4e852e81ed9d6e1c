"""Multi-GPU host logic on CPU: ray sharding + image all-gather with the gloo backend,
world_size 2 (the N>1 path of bench.py / pipeline.py without a GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerfactor_b200.pipeline import shard_range, gather_image


def test_shard_range_partitions_every_ray_once():
    for n in (0, 1, 7, 640000, 640001):
        for world in (1, 2, 3, 8):
            seen = np.zeros(n, np.int32)
            prev_end = 0
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert a == prev_end or a == n
                seen[a:b] += 1
                prev_end = b
            assert np.all(seen == 1)
            # per-rank load differs by at most ceil(n / world)
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0]
                     for r in range(world)]
            assert max(sizes) <= (n + world - 1) // world


def _worker(rank, world, port, n_total, out_q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = shard_range(n_total, rank, world)
    idx = torch.arange(a, b, dtype=torch.float32)
    local = torch.stack((idx, idx * 2, idx * 3 + rank * 0), dim=1)   # "rgb" of my rays
    full = gather_image(local, n_total, rank, world)
    exp = torch.arange(n_total, dtype=torch.float32)
    ok = bool(torch.equal(full, torch.stack((exp, exp * 2, exp * 3), dim=1)))
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out_q.put((rank, ok, float(t.item())))
    dist.destroy_process_group()


def test_gather_image_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    world, n_total = 2, 1001            # ragged: 501 + 500 rays
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(t == float(world) for _, _, t in res)

"""Data-parallel training step on CPU: world_size 2 over gloo, on the test double of the ctypes
layer (tests/cpu_backend.py).  Two ranks with half of the rays each must take the same optimizer
step as one rank with all rays: per-ray loss normalised by the GLOBAL batch
(tf.nn.compute_average_loss, trainvali.py:282-283), gradients summed by ONE all-reduce over the
flat parameter buffer.  The GPU / NCCL version of the same check is tools/dp_check.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _setup():
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cpu_backend
    mpatch = pytest.MonkeyPatch()
    ctx = cpu_backend.install(mpatch)
    return mpatch, ctx


def _make(ctx, seed=7):
    from nerfactor_b200 import synth, config as nfconfig
    from nerfactor_b200.models.nerfactor_microfacet import Model
    params = synth.make_stage_b_params(seed, 'microfacet', light_hw=(2, 4))
    return Model(nfconfig.default_config('nerfactor_microfacet', light_h=2), params=params,
                 ctx=ctx, precision='fp32')


def _data(n=32):
    from nerfactor_b200 import synth
    full = synth.make_stage_b_batch(3, n, 8, fg_frac=1.0)
    noise = (0.01 * np.random.default_rng(5).standard_normal((n, 3))).astype(np.float32)
    return full, noise


def _worker(rank, world, port, q):
    torch.set_num_threads(1)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mpatch, ctx = _setup()
    from nerfactor_b200.trainvali import Trainer
    full, noise = _data()
    n = noise.shape[0]
    a, b = rank * n // world, (rank + 1) * n // world
    shard = tuple(x[a:b] if hasattr(x, '__len__') and len(x) == n else x for x in full)
    tr = Trainer(_make(ctx), world_size=world, rank=rank, precision='fp32')
    losses = [float(tr.train_step(shard, xyz_noise=noise[a:b])) for _ in range(2)]
    q.put((rank, losses, tr.flat.clone().numpy(), tr.iterations))
    mpatch.undo()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_with_all_rays():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mctx = mp.get_context('spawn')
    q = mctx.Queue()
    world = 2
    procs = [mctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, all rays
    mpatch, ctx = _setup()
    try:
        from nerfactor_b200.trainvali import Trainer
        full, noise = _data()
        ref = Trainer(_make(ctx), precision='fp32')
        start = ref.flat.clone().numpy()
        ref_losses = [float(ref.train_step(full, xyz_noise=noise)) for _ in range(2)]
        want = ref.flat.numpy()
    finally:
        mpatch.undo()
    (_, l0, p0, it0), (_, l1, p1, it1) = res
    assert it0 == it1 == 2
    assert np.array_equal(p0, p1)                              # replicas stay in lock-step
    assert np.allclose(l0, l1) and np.allclose(l0, ref_losses, rtol=1e-5)   # global-batch loss
    step = np.abs(want - start).max()
    assert step > 1e-4 and np.abs(p0 - want).max() < 1e-5 * max(1., step / 1e-3)

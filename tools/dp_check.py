"""Data-parallel equivalence (run under torchrun, 2 GPUs): two ranks with half of the rays
each must take the same optimizer step as one rank with all rays (trainvali.py:282-289)."""
import os, sys, json, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor import Model
from nerfactor_b200.trainvali import Trainer
ctx = _lib.Context(local)
params = synth.make_stage_b_params(7, 'learned', light_hw=(16, 32))
def make():
    return Model(nfconfig.default_config('nerfactor'), params=params, ctx=ctx, precision='fp32')
n = 256
full = synth.make_stage_b_batch(3, n, 512, fg_frac=1.0)
noise = (0.01 * np.random.default_rng(5).standard_normal((n, 3))).astype(np.float32)
a, b = rank * n // world, (rank + 1) * n // world
shard = tuple(x[a:b] if hasattr(x, '__len__') and len(x) == n else x for x in full)
tr = Trainer(make(), world_size=world, rank=rank)
loss_dp = tr.train_step(shard, xyz_noise=noise[a:b])
out = {'rank': rank, 'loss_dp': float(loss_dp)}
if rank == 0:
    ref = Trainer(make(), world_size=1)
    loss_1 = ref.train_step(full, xyz_noise=noise)
    d = (tr.flat - ref.flat).abs().max().item()
    out.update(loss_single=float(loss_1), max_param_diff=d,
               max_update=float((ref.flat - Trainer(make()).flat).abs().max()))
    # timing of a reference-size step: 1024 rays x 512 lights per rank
dist.barrier()
big = synth.make_stage_b_batch(9, 1024, 512, fg_frac=1.0)
for _ in range(2):
    tr.train_step(big)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    tr.train_step(big)
torch.cuda.synchronize()
out['train_step_ms_1024rays_per_rank'] = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps(out), flush=True)
dist.destroy_process_group()

"""Ray-sharded single view on N GPUs (torchrun): shards + all_gather == unsharded render."""
import os, sys, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
os.environ['NCCL_DEBUG'] = 'WARN'
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor_microfacet import Model
from nerfactor_b200.models.nerf import Model as NerfModel
from nerfactor_b200.pipeline import ViewRenderer, shard_range, gather_image
ctx = _lib.Context(local)
nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx, precision='f16')
model = Model(nfconfig.default_config('nerfactor_microfacet'),
              params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx)
vr = ViewRenderer(nerf, model, n_samples=64)
h = w = 101                      # 10201 rays: ragged split
n = h * w
a, b = shard_range(n, rank, world)
mine = vr.render(synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w, ray_range=(a, b))['rgb']
full = gather_image(mine.contiguous(), n, rank, world)
ref = vr.render(synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w)['rgb']
out = {'rank': rank, 'shard': [a, b], 'max_abs_diff': float((full - ref).abs().max()),
       'equal': bool(torch.equal(full, ref))}
print(json.dumps(out), flush=True)
dist.destroy_process_group()

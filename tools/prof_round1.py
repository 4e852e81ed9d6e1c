"""Profiling target for the round-1 ncu captures (see profiles/README.md):
  1. one full 800x800 view (Stage A single pass S=128 + Stage B microfacet L=512),
  2. hierarchical Stage A with tcgen05 sigma + d sigma/dx normals on a 200x200 view,
  3. one training Dense layer at step size (524288 rows, 128 -> 128, bf16 operands): forward,
     act-backward + bias gradient, data gradient, weight gradient.
Usage: python tools/prof_round1.py [image side, default 800]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200 import geometry_from_nerf as gfn
from nerfactor_b200.models.nerfactor_microfacet import Model
from nerfactor_b200.models.nerf import Model as NerfModel
from nerfactor_b200.pipeline import ViewRenderer

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 800
ctx = _lib.default_context()
nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx,
                 precision='f16')
model = Model(nfconfig.default_config('nerfactor_microfacet'),
              params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx)
vr = ViewRenderer(nerf, model, n_samples=128)
vr.render(synth.look_at_c2w(), synth.CAM_ANGLE_X, hw, hw)
torch.cuda.synchronize()

ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, 200, 200, normalize=True)
gfn.compute_depth_and_normal(nerf, ro, rd, nfconfig.default_config('nerf'), precision='f16')
torch.cuda.synchronize()

m = 1024 * 512
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn((m, 128), device='cuda', generator=g)
w = torch.randn((128, 128), device='cuda', generator=g) * 0.1
b = torch.zeros((128,), device='cuda')
y = _lib.dense_fwd(ctx, x, None, w, b, 'relu', 'bf16')
dy = torch.randn((m, 128), device='cuda', generator=g)
_lib.dense_bwd(ctx, x, None, w, y, dy, 'relu', True, False, 'bf16')
torch.cuda.synchronize()

"""Short profiling target: one 200x200 view (40k rays, S=128, L=512) through the pipeline."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor_microfacet import Model
from nerfactor_b200.models.nerf import Model as NerfModel
from nerfactor_b200.pipeline import ViewRenderer
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = _lib.default_context()
nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx, precision='f16')
model = Model(nfconfig.default_config('nerfactor_microfacet'), params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx)
vr = ViewRenderer(nerf, model, n_samples=128)
for _ in range(2):
    vr.render(synth.look_at_c2w(), synth.CAM_ANGLE_X, hw, hw)
torch.cuda.synchronize()

"""Times the two big tcgen05 kernels at the full bench size (800x800, S=128, L=512).
Variant knobs: NF_SIGMA_CLUSTER = 1|2|4, NF_SIGMA_PAIR = 1."""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerf import Model as NerfModel
from nerfactor_b200.models.nerfactor_microfacet import Model
ctx = _lib.default_context()
nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx)
model = Model(nfconfig.default_config('nerfactor_microfacet'), params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx)
n, S = 640000, 128
ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, 800, 800, normalize=True)
z = _lib.gen_z(ctx, 2., 6., S, n)
xyz = (ro + rd * 3.0).contiguous()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {'cluster': os.environ.get('NF_SIGMA_CLUSTER'), 'pair': os.environ.get('NF_SIGMA_PAIR')}
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
if which in ('both', 'sigma'):
    out['sigma_ms'] = t(lambda: _lib.sigma_fwd(ctx, nerf.packed_sigma(True), ro, rd, z, None, 'f16'))
if which in ('both', 'lvis'):
    out['lvis_ms'] = t(lambda: model._pred_lvis_at(xyz))
print(json.dumps(out))

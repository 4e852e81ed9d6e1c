import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor import Model
from oracle import stage_b
ctx = _lib.default_context()
params = synth.make_stage_b_params(5, 'learned')
m = Model(nfconfig.default_config('nerfactor'), params=params, ctx=ctx)
om = stage_b.StageB(params, {'brdf': 'learned'})
rng = np.random.default_rng(1)
out = {}
for n in (1, 127, 128, 1000):
    xyz = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    xt = torch.tensor(xyz).cuda()
    for name in ('normal', 'albedo', 'brdf_z'):
        pm = m._packed_mlp(name, 'point', n_freqs_a=10)
        a32 = _lib.point_mlp_fwd(ctx, pm, xt, 1.0, 'fp32').cpu().numpy()
        a3 = _lib.point_mlp_fwd(ctx, pm, xt, 1.0, 'f16x3').cpu().numpy()
        o = om._point_mlp(name, torch.tensor(xyz)).numpy()
        out['%s_n%d' % (name, n)] = [float(np.abs(a3 - a32).max()), float(np.linalg.norm(a3 - o) / np.linalg.norm(o)),
                                     float(np.linalg.norm(a32 - o) / np.linalg.norm(o))]
print(json.dumps(out))
n = 640000
xt = torch.tensor(rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)).cuda()
pm = m._packed_mlp('albedo', 'point', n_freqs_a=10)
for prec in ('fp32', 'f16x3'):
    _lib.point_mlp_fwd(ctx, pm, xt, 1.0, prec); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.point_mlp_fwd(ctx, pm, xt, 1.0, prec)
    e1.record(); torch.cuda.synchronize()
    print(prec, e0.elapsed_time(e1) / 5, 'ms')

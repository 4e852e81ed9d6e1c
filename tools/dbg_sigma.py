"""sigma tcgen05 kernel vs FP32 kernel vs oracle + timing (env NF_SIGMA_CLUSTER = 1|2|4)."""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerf import Model
from oracle import stage_a
ctx = _lib.default_context()
nerf_p = synth.make_nerf_params(3)
m = Model(nfconfig.default_config('nerf'), params=nerf_p, ctx=ctx)
out = {'cluster': os.environ.get('NF_SIGMA_CLUSTER', 'default')}
for (h, w, S) in ((8, 8, 32), (37, 5, 128), (64, 64, 77)):
    ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w, normalize=True)
    z = _lib.gen_z(ctx, 2., 6., S, h * w)
    s32 = _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, None, 'fp32')
    s16 = _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, None, 'f16')
    torch.cuda.synchronize()
    a, b = s32.cpu().numpy(), s16.cpu().numpy()
    key = '%dx%dx%d' % (h, w, S)
    out[key + '_maxabs'] = float(np.abs(a - b).max())
    out[key + '_rel'] = float(np.linalg.norm(a - b) / np.linalg.norm(a))
    out[key + '_mean32'] = float(a.mean()); out[key + '_mean16'] = float(b.mean())
    if (h, w, S) == (8, 8, 32):
        pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3).cpu()
        so = stage_a.eval_sigma_mlp(nerf_p, pts, True).reshape(h * w, S).numpy()
        out['oracle_vs_f16_rel'] = float(np.linalg.norm(so - b) / np.linalg.norm(so))
        out['oracle_vs_f32_rel'] = float(np.linalg.norm(so - a) / np.linalg.norm(so))
        _, occ32, d32, _, _ = _lib.composite(ctx, s32, z, ro, rd)
        _, occ16, d16, _, _ = _lib.composite(ctx, s16, z, ro, rd)
        out['depth_maxabs'] = float((d32 - d16).abs().max())
    # bbox masking
bb = [-1., 1., -1., 1., -1., 1.]
s16b = _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, bb, 'f16')
s32b = _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, bb, 'fp32')
out['bbox_zero_match'] = bool(((s16b == 0) == (s32b == 0)).all().item())
print(json.dumps(out))
# timing at full size
n, S = 640000, 128
ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, 800, 800, normalize=True)
z = _lib.gen_z(ctx, 2., 6., S, n)
for _ in range(2):
    _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, None, 'f16')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    _lib.sigma_fwd(ctx, m.packed_sigma(True), ro, rd, z, None, 'f16')
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
out['full_ms'] = ms; out['full_tflops'] = n * S * 982528 / ms / 1e9
print(json.dumps(out))

"""GPU bring-up diagnostics (run under gpurun): tcgen05 single-tile self-test and
fused-kernel comparisons; writes gpurun_out/bringup.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig   # noqa: E402
from oracle import stage_b, brdf as obrdf                      # noqa: E402

out = {}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
ctx = _lib.default_context()
out['sm_count'] = ctx.sm_count
rng = np.random.default_rng(0)
for K in (16, 32, 128):
    a = rng.integers(-4, 5, (128, K)).astype(np.float32)
    b = rng.integers(-4, 5, (128, K)).astype(np.float32)
    ref = a @ b.T
    for swap in (0,):
        try:
            got = _lib.selftest_umma(ctx, torch.tensor(a).cuda(), torch.tensor(b).cuda(), False)
            torch.cuda.synchronize()
            got = got.cpu().numpy()
            err = float(np.abs(got - ref).max())
            out['umma_K%d_swap%d' % (K, swap)] = err
            if err != 0 and swap == 0:
                np.save(os.path.join(ROOT, 'gpurun_out', 'umma_K%d_got.npy' % K), got)
                np.save(os.path.join(ROOT, 'gpurun_out', 'umma_K%d_ref.npy' % K), ref)
        except Exception as e:  # noqa
            out['umma_K%d_swap%d' % (K, swap)] = 'ERR ' + str(e)
print(json.dumps(out))

# fused lvis kernel: tc vs simt vs oracle
lh, lw = 16, 32
params = synth.make_stage_b_params(5, 'learned', light_hw=(lh, lw))
lxyz, lareas = obrdf.gen_light_xyz(lh, lw)
om = stage_b.StageB(params, {'brdf': 'learned'}, lxyz=lxyz, lareas=lareas)
from nerfactor_b200.models.nerfactor import Model   # noqa: E402
m = Model(nfconfig.default_config('nerfactor'), params=params, ctx=ctx)
n = 300
xyz = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
xt = torch.tensor(xyz).cuda()
lv_o = om.pred_lvis_at(torch.tensor(xyz), om.calc_ldir(torch.tensor(xyz))).numpy()
mlp_l = m._packed_mlp('lvis', 'lvis', n_freqs_a=10, n_freqs_b=4)
for prec in ('fp32', 'f16', 'bf16'):
    try:
        lv = _lib.lvis_fwd(ctx, mlp_l, xt, m.lxyz, 1.0, prec)
        torch.cuda.synchronize()
        lv = lv.cpu().numpy()
        out['lvis_%s_maxabs' % prec] = float(np.abs(lv - lv_o).max())
        out['lvis_%s_rel' % prec] = float(np.linalg.norm(lv - lv_o) / np.linalg.norm(lv_o))
        if prec == 'f16':
            np.save(os.path.join(ROOT, 'gpurun_out', 'lvis_f16.npy'), lv[:8])
            np.save(os.path.join(ROOT, 'gpurun_out', 'lvis_ref.npy'), lv_o[:8])
    except Exception as e:  # noqa
        out['lvis_%s' % prec] = 'ERR ' + str(e)
print(json.dumps(out))

# quick timing of the lvis kernel at a medium size
try:
    n = 148 * 2 * 40
    xt = torch.tensor(rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)).cuda()
    for prec in ('f16', 'fp32'):
        _lib.lvis_fwd(ctx, mlp_l, xt, m.lxyz, 1.0, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _lib.lvis_fwd(ctx, mlp_l, xt, m.lxyz, 1.0, prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out['lvis_%s_ms_n%d' % (prec, n)] = ms
        out['lvis_%s_tflops' % prec] = n * 512 * 144640 / ms / 1e9
except Exception as e:  # noqa
    out['timing'] = 'ERR ' + str(e)
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'bringup.json'), 'w'), indent=1)

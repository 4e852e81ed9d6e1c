"""Times the sigma-network kernel variants on the headline workload (800 x 800 rays x 128 samples)
and reports their error against the FP32 kernel on a slice:  f16 on the 5-slot ring (default),
f16 on the 4-slot ring, f16e (split positional encoding, 4-slot ring), cluster sizes.

    python tools/time_sigma_variants.py           # on a B200 (gpurun); prints one JSON object"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig          # noqa: E402
from nerfactor_b200.models.nerf import Model                        # noqa: E402

FLOP = 982528


def main():
    ctx = _lib.default_context()
    nerf = Model(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx)
    h = w = int(os.environ.get('IMH', '800'))
    S = 128
    ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w, normalize=True)
    z = _lib.gen_z(ctx, 2., 6., S, h * w)
    mlp = nerf.packed_sigma(True)

    def t(prec, reps=3):
        fn = lambda: _lib.sigma_fwd(ctx, mlp, ro, rd, z, None, prec)
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    k = 4096
    s32 = _lib.sigma_fwd(ctx, mlp, ro[:k].contiguous(), rd[:k].contiguous(), z[:k].contiguous(),
                         None, 'fp32')
    out = {}
    for name, prec, env in (('f16_ring5_cl2', 'f16', {}), ('f16e_ring4_cl2', 'f16e', {}),
                            ('f16_ring5_cl1', 'f16', {'NF_SIGMA_CLUSTER': '1'}),
                            ('f16_ring5_cl4', 'f16', {'NF_SIGMA_CLUSTER': '4'}),
                            ('f16_pair', 'f16', {'NF_SIGMA_PAIR': '1'}),
                            ('bf16_ring5_cl2', 'bf16', {})):
        for kk, v in env.items():
            os.environ[kk] = v
        try:
            ms = t(prec)
            s = _lib.sigma_fwd(ctx, mlp, ro[:k].contiguous(), rd[:k].contiguous(),
                               z[:k].contiguous(), None, prec)
            err = float((s - s32).norm() / s32.norm())
            out[name] = {'ms': ms, 'tflops': h * w * S * FLOP / (ms * 1e-3) / 1e12,
                         'rel_l2_vs_fp32_kernel': err}
        except Exception as e:          # a variant that fails must not hide the others
            out[name] = {'error': repr(e)}
        for kk in env:
            del os.environ[kk]
    print(json.dumps(out))


if __name__ == '__main__':
    main()

"""Times Stage B of the headline view (microfacet BRDF, 512 lights, 640 k points) through
Model.render_rgb: visibility network on the front-lit lights only (default when the visibility
tensor is not an output) vs on every light, and checks that the colours agree bit for bit.

    python tools/time_stage_b.py            # on a B200 (gpurun); prints one JSON object"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig          # noqa: E402
from nerfactor_b200.brdf.renderer import gen_light_xyz              # noqa: E402


def main():
    ctx = _lib.default_context()
    out = {}
    for brdf in ('microfacet', 'learned'):
        from nerfactor_b200.models import nerfactor, nerfactor_microfacet
        cls = nerfactor_microfacet.Model if brdf == 'microfacet' else nerfactor.Model
        cfgname = 'nerfactor_microfacet' if brdf == 'microfacet' else 'nerfactor'
        lh, lw = 16, 32
        params = synth.make_stage_b_params(21, brdf, light_hw=(lh, lw))
        m = cls(nfconfig.default_config(cfgname, light_h=lh), params=params, ctx=ctx, precision='f16')
        lxyz, lareas = gen_light_xyz(lh, lw)
        m.set_lights(lxyz.reshape(-1, 3), lareas.reshape(-1))
        m.light_res = (lh, lw)
        n = 640000 if brdf == 'microfacet' else 160000
        batch = synth.make_stage_b_batch(22, n, 1, fg_frac=1.0)

        def t(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        a = m.render_rgb(batch)['rgb']
        b = m.render_rgb(batch, all_lights=True)['rgb']
        # visibility tensor with (all_lights='front_lit') and without the front-lit restriction
        sub = tuple(x[:4096] if hasattr(x, 'shape') and x.shape[:1] == (n,) else x for x in batch)
        lv_all = m.render_rgb(sub, want_lvis=True)['lvis']
        lv_fl = m.render_rgb(sub, want_lvis=True, all_lights='front_lit')['lvis']
        nz = lv_fl != 0
        out[brdf + '_lvis'] = {'front_lit_fraction': float(nz.float().mean()),
                               'max_abs_diff_on_front_lit': float((lv_all - lv_fl)[nz].abs().max())}
        out[brdf] = {'points': n, 'front_lit_ms': t(lambda: m.render_rgb(batch)),
                     'all_lights_ms': t(lambda: m.render_rgb(batch, all_lights=True)),
                     'rgb_max_abs_diff': float((a - b).abs().max())}
    print(json.dumps(out))


if __name__ == '__main__':
    main()

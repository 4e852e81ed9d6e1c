"""Device timing of the rendering-equation kernel variants at the bench size (640 k points x 512
lights): analytic GGX lobe (brdf_kind 0) vs pre-computed lobe (brdf_kind 1: reads lvis + spec),
1 / 4 / 8 env-maps.  Prints one JSON object; run on a B200 (gpurun)."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfactor_b200 import _lib  # noqa: E402
from nerfactor_b200.brdf.renderer import gen_light_xyz  # noqa: E402


def main():
    ctx = _lib.default_context()
    n, lh = 640000, 16
    L = 2 * lh * lh
    g = torch.Generator(device='cuda').manual_seed(0)
    r = lambda *s: torch.rand(*s, device='cuda', generator=g)
    xyz = (r(n, 3) * 2 - 1).contiguous()
    nrm = torch.nn.functional.normalize(xyz + 0.1 * (r(n, 3) - 0.5), dim=1).contiguous()
    cam = torch.tensor([3., 1.7, 2.], device='cuda').expand(n, 3).contiguous()
    alb, lvis, spec, rough = r(n, 3), r(n, L), r(n, L), (0.1 + 0.9 * r(n, 1)).contiguous()
    lxyz, lareas = gen_light_xyz(lh, 2 * lh)
    lxyz = torch.as_tensor(lxyz.reshape(-1, 3).astype(np.float32)).cuda()
    lareas = torch.as_tensor(lareas.reshape(-1).astype(np.float32)).cuda()

    def t(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {'n': n, 'L': L}
    for E in (1, 4, 8):
        light = r(E, L, 3).contiguous()
        ms0 = t(lambda: _lib.integrate_fwd(ctx, xyz, nrm, cam, alb, lvis, lxyz, lareas, light,
                                           rough=rough, f0=0.04))
        ms1 = t(lambda: _lib.integrate_fwd(ctx, xyz, nrm, cam, alb, lvis, lxyz, lareas, light,
                                           spec=spec, spec_scale=1.0))
        b0 = n * (4 * L + 64 + 12 * E)
        b1 = n * (8 * L + 64 + 12 * E)
        out['E%d' % E] = {'ggx_ms': ms0, 'ggx_GBs': b0 / ms0 / 1e6,
                          'spec_ms': ms1, 'spec_GBs': b1 / ms1 / 1e6}
    print(json.dumps(out))


if __name__ == '__main__':
    main()

"""Accuracy of the two rendering-equation kernels against an fp64 restatement on the inputs of
tests/test_gpu_parity.py::test_full_size_properties (linear output, no tone curve)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cpu_backend as cb  # noqa: E402  (fp64 pair terms from the oracle)
from nerfactor_b200 import _lib, synth  # noqa: E402
from nerfactor_b200.brdf.renderer import gen_light_xyz  # noqa: E402


def main():
    ctx = _lib.default_context()
    n, L = 4000, 512
    batch = synth.make_stage_b_batch(32, n, L, fg_frac=1.0)
    lxyz, lareas = gen_light_xyz(16, 32)
    lx = torch.as_tensor(lxyz.reshape(-1, 3).astype(np.float32))
    la = torch.as_tensor(lareas.reshape(-1).astype(np.float32))
    xyz, nrm, cam = [torch.as_tensor(batch[i]) for i in (6, 7, 2)]
    lvis = torch.as_tensor(batch[8])
    out = {}
    for rough_v in (0.4, 0.2, 0.7):
        alb = torch.full((n, 3), .5)
        rough = torch.full((n, 1), rough_v)
        c64 = cb._pair_terms(xyz.double(), nrm.double(), cam.double(), alb.double(), lvis.double(),
                             lx.double(), la.double(), rough.double(), None, 0.04, 1.0)
        truth = (c64.sum(1) * 1e-3).numpy()
        c32 = cb._pair_terms(xyz, nrm, cam, alb, lvis, lx, la, rough, None, 0.04, 1.0)
        o32 = (c32.sum(1) * 1e-3).double().numpy()
        d = lambda t: t.cuda().contiguous()
        args = dict(lxyz=d(lx), lareas=d(la), rough=d(rough), f0=0.04, linear2srgb=False)
        white = torch.full((1, L, 3), 1e-3, device='cuda')
        r = _lib.integrate_fwd(ctx, d(xyz), d(nrm), d(cam), d(alb), d(lvis), light=white, **args)
        olat = _lib.integrate_olat_fwd(ctx, d(xyz), d(nrm), d(cam), d(alb), d(lvis),
                                       olat_inten=1e-3, ambient=0., **args)
        packed = r[:, 0].double().cpu().numpy()
        scalar = olat.double().sum(1).cpu().numpy()
        rl = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        e = np.abs(packed - truth).sum(1) / np.abs(truth).sum(1)
        cosv = (torch.nn.functional.normalize(nrm, dim=1) *
                torch.nn.functional.normalize(cam - xyz, dim=1)).sum(1).numpy()
        worst = np.argsort(-e)[:5]
        out['rough%.1f' % rough_v] = {
            'packed_vs_fp64': rl(packed, truth), 'scalar_olat_vs_fp64': rl(scalar, truth),
            'oracle32_vs_fp64': rl(o32, truth), 'packed_vs_scalar': rl(scalar, packed),
            'worst_points_rel_err': e[worst].tolist(), 'their_cos_v': cosv[worst].tolist(),
            'median_point_rel_err': float(np.median(e))}
    print(json.dumps(out))


if __name__ == '__main__':
    main()

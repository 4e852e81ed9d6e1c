"""One NeRF train step (NerfTrainer: 1024 rays, 64 coarse + 128 fine samples as in nerf.ini, FP32
Dense kernels) timed with CUDA events; prints one JSON object.  bench.py runs this in a
subprocess (context row of the bench line; a failure here cannot touch the headline run)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfactor_b200 import _lib, synth, config as nfconfig  # noqa: E402
from nerfactor_b200.models.nerf import Model  # noqa: E402
from nerfactor_b200.trainvali import make_trainer  # noqa: E402


def main():
    ctx = _lib.default_context()
    cfg = nfconfig.default_config('nerf')
    m = Model(cfg, params=synth.make_nerf_params(0), ctx=ctx, precision='fp32')
    tr = make_trainer(m, precision='fp32')
    n = cfg.getint('DEFAULT', 'n_rays_per_step')
    ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, 32, 32)
    sel = torch.randint(0, 1024, (n,), device=ctx.device)
    batch = (None, None, ro[sel].contiguous(), rd[sel].contiguous(),
             torch.rand((n, 3), device=ctx.device))
    losses = [float(tr.train_step(batch)) for _ in range(3)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        losses.append(float(tr.train_step(batch)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_c, n_f = cfg.getint('DEFAULT', 'n_samples_coarse'), cfg.getint('DEFAULT', 'n_samples_fine')
    samples = n * (n_c + n_c + n_f)
    flop = 3 * samples * (982528 + 2 * (256 * 256 + 283 * 128 + 128 * 3))     # fwd + dgrad + wgrad
    print(json.dumps({'what': 'NerfTrainer.train_step, %d rays, %d coarse + %d fine-pass samples, '
                              'FP32 Dense kernels' % (n, n_c, n_c + n_f),
                      'ms': ms, 'rays_per_s': n / (ms * 1e-3), 'tflops': flop / (ms * 1e-3) / 1e12,
                      'loss_first': losses[0], 'loss_last': losses[-1],
                      'finite': bool(np.isfinite(losses).all())}))


if __name__ == '__main__':
    main()

"""Prints the hand-over timeline of the sigma kernel (csrc/nf_sigma_tc.cu) for one tile of CTA 0:
clock64 stamps written by the issuer thread and by one epilogue thread when NF_SIGMA_DBG=<file>
is set (debug plumbing, one extra cudaMalloc + sync per call).

    python tools/sigma_timeline.py [f16|f16e]          # on a B200 (gpurun)

Columns, clocks relative to the first stamp, per (layer, N-half):
  commit  issuer: tcgen05.commit of that accumulator half issued (NOT completed)
  wait_a  issuer: returned from waiting for the activations of K-block <half> of THIS layer
  wake    epilogue: woke up on "accumulator half complete"
  ld      epilogue: tcgen05.ld + wait::ld done
  st      epilogue: converted, tcgen05.st + wait::st done
  arr     epilogue: arrived on "activations ready" """
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig          # noqa: E402
from nerfactor_b200.models.nerf import Model                        # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else 'f16'
    ctx = _lib.default_context()
    nerf = Model(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx)
    h = w = 400
    ro, rd = _lib.gen_rays(ctx, synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w, normalize=True)
    z = _lib.gen_z(ctx, 2., 6., 128, h * w)
    mlp = nerf.packed_sigma(True)
    _lib.sigma_fwd(ctx, mlp, ro, rd, z, None, prec)
    torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), 'dbg.txt')
    os.environ['NF_SIGMA_DBG'] = path
    _lib.sigma_fwd(ctx, mlp, ro, rd, z, None, prec)
    torch.cuda.synchronize()
    del os.environ['NF_SIGMA_DBG']
    rows = [[int(v) for v in l.split()] for l in open(path)]
    t0 = min(v for r in rows for v in r if v > 0)
    out = []
    print('layer half  commit  wait_a    wake      ld      st     arr')
    for i, r in enumerate(rows):
        rel = [v - t0 if v > 0 else -1 for v in r]
        out.append(rel)
        print('%5d %4d %7d %7d %7d %7d %7d %7d' % (i // 2, i % 2, rel[0], rel[1], rel[2], rel[3], rel[4], rel[5]))
    print(json.dumps({'precision': prec, 'stamps': out}))


if __name__ == '__main__':
    main()

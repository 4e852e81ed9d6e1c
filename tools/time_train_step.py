"""configs[3]-sized train step (1024 rays x 512 lights, jitter on): ms per step with the bf16
tensor-core Dense kernels (CUDA-graph replay) for both model families.  One JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig            # noqa: E402
from nerfactor_b200.trainvali import Trainer                            # noqa: E402


def main():
    ctx = _lib.default_context()
    out = {}
    for kind in ('learned', 'microfacet'):
        name = 'nerfactor' if kind == 'learned' else 'nerfactor_microfacet'
        from importlib import import_module
        Model = import_module('nerfactor_b200.models.' + name).Model
        m = Model(nfconfig.default_config(name), params=synth.make_stage_b_params(0, kind), ctx=ctx,
                  precision='fp32')
        tb = synth.make_stage_b_batch(2, 1024, 512, fg_frac=1.0)
        for prec in ('bf16', 'fp32'):
            tr = Trainer(m, precision=prec)
            for _ in range(3):
                tr.train_step(tb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                loss = tr.train_step(tb)
            e1.record()
            torch.cuda.synchronize()
            out['%s_%s' % (kind, prec)] = {'ms': e0.elapsed_time(e1) / 5, 'loss': float(loss)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()

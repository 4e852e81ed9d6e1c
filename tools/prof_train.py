"""Launch list target: one configs[3]-sized train step (1024 rays x 512 lights, jitter on, bf16
tensor-core Dense kernels), eager so that every kernel is visible to ncu.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_train.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig            # noqa: E402
from nerfactor_b200.trainvali import Trainer                            # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'learned'
ctx = _lib.default_context()
if kind == 'learned':
    from nerfactor_b200.models.nerfactor import Model
    m = Model(nfconfig.default_config('nerfactor'), params=synth.make_stage_b_params(0, 'learned'),
              ctx=ctx, precision='fp32')
else:
    from nerfactor_b200.models.nerfactor_microfacet import Model
    m = Model(nfconfig.default_config('nerfactor_microfacet'),
              params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx, precision='fp32')
tb = synth.make_stage_b_batch(2, 1024, 512, fg_frac=1.0)
tr = Trainer(m, precision='bf16')
for _ in range(2):
    tr.train_step(tb, graph=False)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('step')
tr.train_step(tb, graph=False)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()

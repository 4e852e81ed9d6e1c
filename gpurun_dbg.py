import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor import Model as LearnedModel
from nerfactor_b200.trainvali import Trainer
ctx = _lib.default_context()
lm2 = LearnedModel(nfconfig.default_config('nerfactor'), params=synth.make_stage_b_params(0, 'learned'), ctx=ctx, precision='fp32')
tb = synth.make_stage_b_batch(2, 1024, 512, fg_frac=1.0)
for prec in ('fp32', 'bf16'):
    tr = Trainer(lm2, precision=prec)
    for _ in range(3): tr.train_step(tb)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(5): tr.train_step(tb)
    torch.cuda.synchronize()
    print(prec, 'wall ms/step', (time.time() - t) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): tr.train_step(tb)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))

"""Profiling target: one 24576-point chunk of the learned model's fused Stage B (512 lights): the
front-lit visibility kernel lvis_tc3_kernel<0, 1> and the learned-BRDF kernel brdf_tc3_kernel<0>.
    ncu --set full -k regex:tc3_kernel -c 2 python tools/prof_tc3.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig          # noqa: E402
from nerfactor_b200.models.nerfactor import Model                   # noqa: E402
from nerfactor_b200.brdf.renderer import gen_light_xyz              # noqa: E402

ctx = _lib.default_context()
lh, lw = 16, 32
m = Model(nfconfig.default_config('nerfactor', light_h=lh), params=synth.make_stage_b_params(21, 'learned', light_hw=(lh, lw)),
          ctx=ctx, precision='f16')
lxyz, lareas = gen_light_xyz(lh, lw)
m.set_lights(lxyz.reshape(-1, 3), lareas.reshape(-1))
m.light_res = (lh, lw)
batch = synth.make_stage_b_batch(22, 24576, 1, fg_frac=1.0)
m.render_rgb(batch)
torch.cuda.synchronize()

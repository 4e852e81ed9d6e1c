"""Profiling target for the round-2 ncu captures (profiles/README.md): two full 800 x 800 views
(Stage A single pass S = 128 in the benchmarked 'f16e' mode + Stage B microfacet L = 512) --
once through Model.call (separate full-size light-visibility and rendering-equation kernels: the
per-kernel roofline rows) and once through the fused Stage-B op over point chunks (what bench.py
times).   python tools/prof_round2.py [image side]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib, synth, config as nfconfig                  # noqa: E402
from nerfactor_b200.models.nerfactor_microfacet import Model                # noqa: E402
from nerfactor_b200.models.nerf import Model as NerfModel                   # noqa: E402
from nerfactor_b200.pipeline import ViewRenderer                            # noqa: E402

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 800
ctx = _lib.default_context()
nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx,
                 precision='f16e')
model = Model(nfconfig.default_config('nerfactor_microfacet'),
              params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx)
vr = ViewRenderer(nerf, model, n_samples=128)
# separate full-size kernels (the roofline rows of bench.py: one launch of each hot kernel), then
# the default path (nf_stageB_fused_fwd over L2-resident point chunks)
for fused in (False, True):
    vr.render(synth.look_at_c2w(), synth.CAM_ANGLE_X, hw, hw, fused=fused)
    torch.cuda.synchronize()

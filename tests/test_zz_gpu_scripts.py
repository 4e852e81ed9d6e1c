"""The three drop-in scripts end to end on the real kernels (B200): the scenario of
tests/e2e_scenario.py -- Stage A buffers from a NeRF checkpoint, a short joint optimisation with
bf16 tensor-core Dense kernels + CUDA-graph replay, checkpoint resume, relighting of the test
views with probes / OLAT, visualisation tree and video.  (Named zz so it runs after the kernel
parity tests.)"""
import pytest

import e2e_scenario

pytestmark = pytest.mark.gpu


def test_scripts_end_to_end_on_gpu(tmp_path):
    import torch
    from nerfactor_b200 import _lib
    l0 = _lib.default_context().launches
    e2e_scenario.run(tmp_path, imh=8, light_h=4, n_samples=8, epochs=2, n_rays=64,
                     infer_precision='f16')
    torch.cuda.synchronize()
    assert _lib.default_context().launches - l0 > 100       # the CUDA library did the work

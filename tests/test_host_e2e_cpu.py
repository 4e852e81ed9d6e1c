"""Host logic of the drop-in scripts end to end on the CPU test double (tests/cpu_backend.py):
file formats, dataset -> batch -> Model.call glue, loss, optimizer bookkeeping, checkpoint
naming / resume, visualisation tree.  The same scenario runs on the real kernels in
tests/test_zz_gpu_scripts.py."""
import numpy as np
import pytest
import torch

import cpu_backend
import e2e_scenario
from nerfactor_b200 import config as nfconfig, synth


def test_scripts_end_to_end_on_test_double(tmp_path, monkeypatch):
    cpu_backend.install(monkeypatch)
    e2e_scenario.run(tmp_path, imh=8, light_h=2, n_samples=8, epochs=2, n_rays=32)


def test_model_call_glue_matches_oracle_on_test_double(monkeypatch):
    """Model.call's compaction / scatter / override glue vs the oracle's Model.call on a batch
    with background rays (the kernels themselves are replaced, so this isolates the host code)."""
    from oracle import stage_b, brdf as obrdf
    ctx = cpu_backend.install(monkeypatch)
    from nerfactor_b200.models.nerfactor_microfacet import Model
    lh, lw = 2, 8
    params = synth.make_stage_b_params(7, 'microfacet', light_hw=(lh, lw))
    m = Model(nfconfig.default_config('nerfactor_microfacet', light_h=lh), params=params, ctx=ctx)
    lxyz, lareas = obrdf.gen_light_xyz(lh, lw)
    m.set_lights(lxyz.reshape(-1, 3), lareas.reshape(-1))
    m.light_res = (lh, lw)
    batch = synth.make_stage_b_batch(11, 40, lh * lw, fg_frac=0.6)
    probes = synth.make_probes(3, 2, (lh, lw))
    for i, p in enumerate(probes):
        m.novel_probes['p%d' % i] = torch.as_tensor(p)
    pred, gt, lk, to_vis = m.call(batch, 'test', relight_probes=True, relight_olat=True,
                                  albedo_scales=np.array([0.5, 1., 2.], np.float32))
    om = stage_b.StageB(params, {'brdf': 'microfacet'}, lxyz=lxyz, lareas=lareas)
    olat = om.novel_olat((lh, lw))                      # the 16 one-hot env-maps, (i, j) order
    opred, _, _ = om.call(batch, 'test', relight_lights=list(probes) + olat,
                          albedo_scales=np.array([0.5, 1., 2.], np.float32))
    want = {k: opred[k].numpy() for k in ('rgb', 'normal', 'lvis', 'albedo', 'brdf')}
    want['rgb_probes'] = opred['rgb_relit'].numpy()[:, :2]
    want['rgb_olat'] = opred['rgb_relit'].numpy()[:, 2:]
    for k, b in want.items():
        a = pred[k].numpy()
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() < 2e-5, (k, np.abs(a - b).max())
    bg = batch[5][:, 0] == 0
    assert bg.any() and np.all(pred['rgb'].numpy()[bg] == 0)           # nerfactor.py:268-293
    assert to_vis['pred_rgb'] is pred['rgb'] and set(lk) == {
        'mode', 'normal_jitter', 'lvis_jitter', 'brdf_prop_jitter', 'albedo_jitter'}
    with pytest.raises(ValueError):
        m.call(batch, 'predict')

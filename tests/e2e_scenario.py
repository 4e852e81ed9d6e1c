"""The three drop-in scripts run back to back on a tiny synthetic scene written to disk in the
reference's layout -- shared by the CPU host-logic test (tests/test_host_e2e_cpu.py, on the
`tests/cpu_backend.py` test double) and the GPU test (tests/test_zz_gpu_scripts.py, real kernels):

  geometry_from_nerf.main   trained-NeRF checkpoint -> alpha.png / xyz.npy / normal.npy / lvis.npy
  trainvali.main            nerfactor_microfacet, a few epochs, checkpoints + validation vis
  test.main                 relight every test view (probes; OLAT on the last), vis + video
"""
import os
from os.path import exists, join

import numpy as np

from nerfactor_b200 import config as nfconfig, synth
from nerfactor_b200.util import io as ioutil, tfckpt


def run(tmp, imh=8, light_h=2, n_samples=8, epochs=2, n_rays=32, train_precision=None,
        infer_precision='f16'):
    from nerfactor_b200 import geometry_from_nerf as gfn, trainvali, test as nftest
    tmp = str(tmp)
    data, surf, env = join(tmp, 'data'), join(tmp, 'surf'), join(tmp, 'envmaps')
    ids = synth.write_scene(data, imh=imh, imw=imh, n_train=2, n_val=1, n_test=2,
                            envmap_dir=env, n_probes=2, light_hw=(light_h, 2 * light_h))
    # ---- a "trained" NeRF: random-init weights saved as a reference-named checkpoint
    nerf_dir = join(tmp, 'out', 'nerf', 'lr1e-4')
    nerf_cfg = nfconfig.default_config(
        'nerf', data_root=data, imh=imh, n_samples_coarse=n_samples, n_samples_fine=n_samples,
        outroot=join(tmp, 'out', 'nerf'))
    ioutil.write_config(nerf_cfg, nerf_dir + '.ini')
    params = synth.make_nerf_params(0)
    tfckpt.write_checkpoint(join(nerf_dir, 'checkpoints', 'ckpt-1'),
                            tfckpt.tensors_from_params(params, step=1))
    # ---- Stage A
    done = gfn.main(['--trained_nerf', nerf_dir, '--out_root', surf, '--light_h', str(light_h),
                     '--imh', str(imh), '--precision', infer_precision])
    assert sorted(done) == sorted(ids)
    L = 2 * light_h * light_h
    for id_ in ids:
        lv = np.load(join(surf, id_, 'lvis.npy'))
        assert lv.shape == (imh, imh, L) and lv.min() >= 0 and lv.max() <= 1
        assert np.load(join(surf, id_, 'normal.npy')).shape == (imh, imh, 3)
        assert exists(join(surf, id_, 'alpha.png')) and exists(join(surf, id_, 'xyz.png'))
    assert gfn.main(['--trained_nerf', nerf_dir, '--out_root', surf, '--light_h', str(light_h),
                     '--imh', str(imh), '--precision', infer_precision]) == []   # all skipped
    # ---- joint optimisation (trainvali.py), resumed once
    cfg = nfconfig.default_config(
        'nerfactor_microfacet', data_root=data, data_nerf_root=surf, imh=imh, light_h=light_h,
        shape_mode='scratch', n_rays_per_step=n_rays, epochs=epochs, ckpt_period=1,
        vali_period=epochs, vali_batches=1, keep_recent_epochs=1, test_envmap_dir=env,
        outroot=join(tmp, 'out', 'nerfactor'), use_nerf_alpha='True')
    ini = join(tmp, 'nerfactor_microfacet.ini')
    ioutil.write_config(cfg, ini)
    argv = ['--config', ini] + (['--precision', train_precision] if train_precision else [])
    outdir = trainvali.main(argv)
    ckptdir = join(outdir, 'checkpoints')
    ckpt = ioutil.latest_checkpoint(ckptdir)
    assert ckpt.endswith('ckpt-%d' % epochs) and exists(ckpt + '.index')
    assert not exists(join(ckptdir, 'ckpt-%d.index' % (epochs - 1)))          # max_to_keep = 1
    assert exists(outdir.rstrip('/') + '.ini')
    vdir = join(outdir, 'vis_vali', 'epoch%09d' % epochs)
    assert exists(join(vdir, 'all.html')) and exists(join(vdir, 'pred_light.png'))
    meta = ioutil.read_json(join(vdir, 'batch000000000', 'metadata.json'))
    assert meta['id'] == 'val_000' and np.isfinite(meta['psnr'])
    summ = [l for l in open(join(outdir, 'summary_train.jsonl'))]
    assert len(summ) == epochs
    # one more epoch from the checkpoint: the step counter continues
    trainvali.main(argv + ['--config_override', 'epochs=%d' % (epochs + 1)])
    ckpt = ioutil.latest_checkpoint(ckptdir)
    assert ckpt.endswith('ckpt-%d' % (epochs + 1))
    assert int(tfckpt.read_checkpoint(ckpt)[
        'optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE']) == 2 * (epochs + 1)     # 2 train views
    # ---- relighting the test views
    outroot, view_at = nftest.main(['--ckpt', ckpt, '--precision', infer_precision])
    assert outroot == join(outdir, 'vis_test', 'ckpt-%d' % (epochs + 1))
    b0, b1 = join(outroot, 'batch000000000'), join(outroot, 'batch000000001')
    for b in (b0, b1):
        for f in ('pred_rgb.png', 'pred_albedo.png', 'pred_brdf.png', 'pred_normal.png',
                  'pred_lvis.png', 'pred_rgb_probes_probe0.png', 'pred_rgb_probes_probe1.png',
                  'metadata.json'):
            assert exists(join(b, f)), (b, f)
    olat = [f for f in os.listdir(b1) if f.startswith('pred_rgb_olat_')]
    assert len(olat) == L // 2 and not any(f.startswith('pred_rgb_olat_') for f in os.listdir(b0))
    assert view_at.endswith('.mp4') and os.path.getsize(view_at) > 0
    # an albedo edit goes to its own directory (test.py:142-146)
    outroot2, _ = nftest.main(['--ckpt', ckpt, '--tgt_albedo', 'gold', '--no_video',
                               '--precision', infer_precision])
    assert outroot2 == outroot + '_gold' and exists(join(outroot2, 'batch000000000', 'pred_rgb.png'))
    from nerfactor_b200.util import img as imgutil
    alb = imgutil.read(join(outroot2, 'batch000000000', 'pred_albedo.png')).reshape(-1, 3)
    fg = imgutil.read(join(surf, 'test_000', 'alpha.png')).reshape(-1) >= 0.8 * 255
    if fg.any():          # gold = (1, .843, 0), gamma 1/2.2, truncated to 8 bit
        want = (np.array([1., 0.843, 0.]) ** (1 / 2.2) * 255).astype(np.uint8)
        assert np.abs(alb[fg].astype(int) - want.astype(int)).max() <= 1
    return outdir

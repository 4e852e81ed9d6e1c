"""CUDA kernels against outputs of the REFERENCE'S OWN CODE: the unmodified model files of
/root/reference run op by op through the TensorFlow shim (tests/golden/tfshim) in the build
container by tests/golden/make_golden_tfshim.py -> tests/golden/ref_tfshim_*.npz.  Same
tolerances as the oracle-golden tests of test_gpu_parity.py (the oracle reproduces these fixtures
to <= 1e-6, tests/test_oracle_pinning.py)."""
import os

import numpy as np
import pytest
import torch

from nerfactor_b200 import config as nfconfig, synth
from test_gpu_parity import rel_l2, dev, _stage_b, _nerf_model, ctx  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('brdf', ['microfacet', 'learned'])
@pytest.mark.parametrize('precision', ['fp32', 'f16'])
def test_model_call_vs_reference_code_via_shim(ctx, golden_dir, brdf, precision):
    """Model.call on the GPU against outputs of the REFERENCE'S OWN model code (unmodified files of
    /root/reference run op by op through tests/golden/tfshim by make_golden_tfshim.py): the
    reference's native light grid light_h = 4 (32 lights), 80 rays with background rows, probe and
    OLAT relighting, the test.py edits."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_b_%s.npz' % brdf))
    lh = int(g['light_h'])
    m, _, _ = _stage_b(ctx, brdf, lh, 2 * lh, int(g['seed_params']), precision)
    batch = synth.make_stage_b_batch(int(g['seed_batch']), int(g['n_rays']), 2 * lh * lh)
    for i, p in enumerate(g['probes']):
        m.novel_probes['p%d' % i] = p
    pred, _, _, _ = m.call(batch, 'test', relight_probes=True, relight_olat=True)
    tol_net = 1e-5 if precision == 'fp32' else 3e-3
    tol_rgb = 1e-5 if precision == 'fp32' else 1e-4      # north-star bar on RGB
    for k in ('normal', 'albedo', 'brdf'):
        assert rel_l2(pred[k].cpu(), g['test_' + k]) < 1e-5, k
    assert rel_l2(pred['lvis'].cpu(), g['test_lvis']) < tol_net
    assert rel_l2(pred['rgb'].cpu(), g['test_rgb']) < tol_rgb
    assert rel_l2(pred['rgb_probes'].cpu(), g['test_rgb_probes']) < tol_rgb
    assert rel_l2(pred['rgb_olat'].cpu(), g['test_rgb_olat']) < tol_rgb
    kw = {'albedo_override': np.array([0.3, 0.5, 0.7], np.float32)}
    if brdf != 'microfacet':
        kw['brdf_z_override'] = np.array([0.01, -0.02, 0.005], np.float32)
    assert rel_l2(m.call(batch, 'test', **kw)[0]['rgb'].cpu(), g['edit_rgb']) < tol_rgb
    p2 = m.call(batch, 'test', albedo_scales=np.array([0.5, 1., 2.], np.float32))[0]
    assert rel_l2(p2['rgb'].cpu(), g['scaled_rgb']) < tol_rgb
    if precision == 'fp32':      # per-ray training loss with the reference's recorded jitter
        pr, gt, lk, _ = m.call(batch, 'train', xyz_noise=g['xyz_noise'])
        loss = m.compute_loss(pr, gt, **lk)
        assert np.allclose(loss.cpu().numpy(), g['train_loss'], atol=2e-6, rtol=1e-4)


@pytest.mark.parametrize('brdf', ['microfacet', 'learned'])
def test_model_call_512_lights_vs_reference_code_via_shim(ctx, golden_dir, brdf):
    """The north-star light configuration (16 x 32 = 512 lights) in the tcgen05 f16 path against
    the reference's own code through the shim: RGB, probe and OLAT relighting within 1e-4."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_b_%s_L512.npz' % brdf))
    lh = int(g['light_h'])
    m, _, _ = _stage_b(ctx, brdf, lh, 2 * lh, int(g['seed_params']), 'f16')
    batch = synth.make_stage_b_batch(int(g['seed_batch']), int(g['n_rays']), 2 * lh * lh)
    for i, p in enumerate(g['probes']):
        m.novel_probes['p%d' % i] = p
    pred, _, _, _ = m.call(batch, 'test', relight_probes=True, relight_olat=True)
    for k in ('normal', 'albedo', 'brdf'):
        assert rel_l2(pred[k].cpu(), g['test_' + k]) < 1e-5, k
    assert rel_l2(pred['lvis'].cpu(), g['test_lvis']) < 3e-3
    assert rel_l2(pred['rgb'].cpu(), g['test_rgb']) < 1e-4
    assert rel_l2(pred['rgb_probes'].cpu(), g['test_rgb_probes']) < 1e-4
    assert rel_l2(pred['rgb_olat'].cpu().numpy()[:, ::16], g['test_rgb_olat']) < 1e-4


def test_stage_a_vs_reference_code_via_shim(ctx, golden_dir):
    """compute_depth_and_normal / compute_light_visibility / eval_sigma_mlp (FP32 kernels) against
    the reference's geometry_from_nerf.py run through the shim (16 coarse + 88 fine samples)."""
    from nerfactor_b200 import geometry_from_nerf as gfn, _lib
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_a.npz'))
    model = _nerf_model(ctx, int(g['seed_nerf']))
    ro, rdn = dev(g['rayo'], ctx), dev(g['rayd_n'], ctx)
    cfg = nfconfig.default_config('nerf', n_samples_coarse=-48, n_samples_fine=8)
    occu, depth, normal = gfn.compute_depth_and_normal(model, ro, rdn, cfg, precision='fp32')
    d = np.abs(depth.cpu().numpy() - g['depth'])
    assert np.median(d) < 1e-4 and np.quantile(d, 0.95) < 5e-3
    assert np.abs(occu.cpu().numpy() - g['occu']).max() < 1e-3
    dn = np.abs(normal.cpu().numpy() - g['normal']).max(axis=1)
    assert np.median(dn) < 1e-3 and np.quantile(dn, 0.95) < 2e-2
    surf = ro + rdn * dev(g['depth'], ctx)[:, None]
    model.precision = 'fp32'
    lv = gfn.compute_light_visibility(model, surf.contiguous(), dev(g['normal'], ctx), cfg,
                                      light_h=int(g['light_h']))
    assert np.allclose(lv.cpu().numpy(), g['lvis_hit'], atol=3e-4)
    pts = dev(g['sigma_pts'], ctx)
    z1 = torch.ones((pts.shape[0], 1), device=ctx.device)
    zero = torch.zeros_like(pts)                         # samples o + z d with d = 0: the points
    for fine, key in ((False, 'sigma_coarse'), (True, 'sigma_fine')):
        sg = _lib.sigma_fwd(ctx, model.packed_sigma(fine), pts, zero, z1, None, 'fp32')
        assert rel_l2(sg.cpu(), g[key]) < 2e-5


def test_nerf_trainer_vs_reference_train_step(ctx, golden_dir):
    """NerfTrainer on the GPU (FP32 Dense kernels) against the reference's NeRF train step run
    through the shim with its four random draws recorded (ref_tfshim_nerf_train_grad.npz):
    renderings, per-ray loss, all 48 gradient tensors; then a few optimizer steps."""
    from nerfactor_b200.models.nerf import Model
    from nerfactor_b200.trainvali import make_trainer
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_nerf_train_grad.npz'))
    cfg = nfconfig.default_config('nerf', n_samples_coarse=int(g['n_c']),
                                  n_samples_fine=int(g['n_f']), perturb=True,
                                  noise_std=float(g['noise_std']))
    m = Model(cfg, params=synth.make_nerf_params(int(g['seed_nerf'])), ctx=ctx, precision='fp32')
    tr = make_trainer(m, precision='fp32')
    batch = (None, None, g['rayo'], g['rayd'], g['rgb'])
    draws = dict(perturb_u=g['perturb_u'], fine_u=g['fine_u'],
                 sigma_noise=(g['noise_coarse'], g['noise_fine']))
    with torch.no_grad():
        _, pred = tr.forward(tr.flat, batch, 'train', **draws)
        z_mine = tr.last_z_all.cpu().numpy()
    assert np.abs(pred['coarse'].cpu().numpy() - g['pred_coarse']).max() < 1e-5
    # own importance sampling: the inverse-CDF lookup is discontinuous in the coarse weights, so
    # single samples may land in a neighbouring bin; the bulk must coincide with the reference's
    assert np.median(np.abs(z_mine - g['z_all'])) < 1e-5
    d = np.abs(pred['fine'].cpu().numpy() - g['pred_fine'])
    assert d.max() < 5e-4 and np.median(d) < 1e-5
    # gradients: replay the reference's recorded samples (they carry no gradient, nerf.py:143), so
    # the lookup's discontinuity is out of the comparison and the bound can be tight
    replay = dict(draws, z_all=g['z_all'])
    with torch.no_grad():
        _, pred = tr.forward(tr.flat, batch, 'train', **replay)
    dfine = np.abs(pred['fine'].cpu().numpy() - g['pred_fine']).max()
    print('nerf trainer: fine rendering with replayed samples, max err %.2e' % dfine)
    assert dfine < 2e-4          # FP32 FFMA kernels vs torch-CPU summation order, 14 samples / ray
    loss, grad = tr.loss_and_grad(batch, **replay)
    assert np.allclose(loss.cpu().numpy(), g['per_example_loss'], atol=1e-4, rtol=1e-3)
    gv = tr.views(grad)
    keys = [k for k in g.files if k.startswith('grad/')]
    assert len(keys) == len(gv) == 48
    # A pre-activation within rounding of zero takes a different ReLU branch under a different
    # summation order (FP32 FFMA kernels vs torch-CPU): that unit's derivative flips for one of the
    # 168 sample rows, which moves a bias gradient by ~1 / 168 of its size (measured: 5.7e-3 of the
    # tensor maximum on fine_enc/5/bias, 2.4e-3 on kernels).  So: a tight bound on the bulk
    # (relative L2 per tensor) and a looser one on the single worst entry.
    worst_l2 = worst_max = 0.
    for k in keys:
        _, net, li, kind = k.split('/')
        want = g[k].astype(np.float32)
        got = gv[(net, int(li), kind)].cpu().numpy()
        e_l2 = rel_l2(got, want)
        e_max = np.abs(got - want).max() / max(np.abs(want).max(), 1e-8)
        worst_l2, worst_max = max(worst_l2, e_l2), max(worst_max, e_max)
        assert e_l2 <= 1e-2 and e_max <= 3e-2, (k, e_l2, e_max)
    print('nerf trainer: worst gradient rel-L2 %.2e, worst entry / tensor max %.2e'
          % (worst_l2, worst_max))
    l0 = float(tr.train_step(batch, **draws))
    for _ in range(5):
        l1 = float(tr.train_step(batch, **draws))
    assert np.isfinite(l0) and l1 < l0


def test_brdf_prior_trainer_vs_reference_train_step(ctx, golden_dir):
    """BrdfTrainer on the GPU (FP32 Dense kernels) against the reference's BRDF-prior train step
    through the shim (ref_tfshim_brdf_train_grad.npz): predictions, per-row loss, the 10 Dense
    gradients and the latent-code gradient."""
    from nerfactor_b200.models.brdf import Model
    from nerfactor_b200.trainvali import make_trainer
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_brdf_train_grad.npz'))
    names = [str(x) for x in g['names']]
    m = Model(nfconfig.default_config('brdf', lr=1e-3), params=synth.make_stage_b_params(5, 'learned'),
              brdf_names=names)
    m.latent_code.z = g['z0']
    i = int(g['i'])
    batch = (names[i], i, 16, 128, 1, g['rusink'], g['refl'])
    pred, gt, lk, _ = m.call(batch, 'vali')
    assert np.abs(pred['brdf'].cpu().numpy() - g['pred_brdf']).max() < 1e-5
    assert np.abs(pred['brdf_reci'].cpu().numpy() - g['pred_brdf_reci']).max() < 1e-5
    tr = make_trainer(m)
    loss, grad = tr.loss_and_grad(batch)
    assert np.allclose(loss.cpu().numpy(), g['per_example_loss'], atol=1e-5, rtol=1e-4)
    gv = tr.views(grad)
    for k in [k for k in g.files if k.startswith('grad/')]:
        parts = k.split('/')
        key = ('z', 0, 'z') if parts[1] == 'z' else (parts[1], int(parts[2]), parts[3])
        want = g[k]
        assert np.abs(gv[key].cpu().numpy() - want).max() <= 2e-3 * max(np.abs(want).max(), 1e-8), k
    l0 = float(tr.train_step(batch))
    for _ in range(30):
        l1 = float(tr.train_step(batch))
    assert np.isfinite(l0) and l1 < l0

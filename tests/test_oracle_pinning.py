"""Pins the oracle against the parts of the reference that run without
TensorFlow (fixtures: tests/golden/ref_pinned.npz, made by make_golden.py from
/root/reference) and checks the oracle against its own frozen goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import brdf as obrdf, stage_b, stage_a
from nerfactor_b200 import synth


def _rel(a, b):
    a = a.detach().numpy() if hasattr(a, 'detach') else np.asarray(a)
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _ref(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_pinned.npz'))


def test_gen_light_xyz_matches_reference(golden_dir):
    ref = _ref(golden_dir)
    for h, w in ((16, 32), (2, 8), (16, 64)):
        xyz, areas = obrdf.gen_light_xyz(h, w)
        assert np.array_equal(xyz, ref['lxyz_%dx%d' % (h, w)])      # bit-exact fp64
        assert np.array_equal(areas, ref['lareas_%dx%d' % (h, w)])
        assert abs(areas.sum() - 4 * np.pi) < 1e-12


def test_sph2cart_matches_reference(golden_dir):
    ref = _ref(golden_dir)
    assert np.array_equal(obrdf.sph2cart(ref['sph_in']), ref['sph_out'])


def test_dir2rusink_matches_reference_numpy_twin(golden_dir):
    ref = _ref(golden_dir)
    a = torch.tensor(ref['rusink_a'])            # fp64 like the NumPy twin
    b = torch.tensor(ref['rusink_b'])
    out = obrdf.dir2rusink(a, b).numpy()
    # the TF version adds eps=1e-6 inside the normalisations; on unit-scale
    # inputs that is invisible at 1e-9
    assert np.allclose(out, ref['rusink_out'], atol=1e-9)


def test_linear2srgb_matches_reference(golden_dir):
    ref = _ref(golden_dir)
    out = stage_b.linear2srgb(torch.tensor(ref['srgb_in'])).numpy()
    assert np.allclose(out, ref['srgb_out'], atol=1e-14)


def test_gen_world2local_matches_reference_numpy_twin(golden_dir):
    ref = _ref(golden_dir)
    out = obrdf.gen_world2local(torch.tensor(ref['w2l_normal'])).numpy()
    # NumPy twin has z=(0,0,1) exactly, the TF one z=(0,0,1)+1e-6 (geom.py:128):
    # the tangent moves by ~1e-6 / |n x z|
    sin_nz = np.linalg.norm(np.cross(ref['w2l_normal'], [0., 0., 1.]), axis=1)
    err = np.abs(out - ref['w2l_out']).max(axis=(1, 2))
    assert np.all(err <= 4e-6 / sin_nz + 1e-7)
    # rows are orthonormal and the last row is the normal
    eye = np.einsum('nij,nkj->nik', out, out)
    assert np.allclose(eye, np.eye(3)[None], atol=1e-6)
    assert np.allclose(out[:, 2, :], ref['w2l_normal'], atol=1e-6)


def test_rendering_equation_matches_reference_sphere_renderer(golden_dir):
    """oracle StageB.calc_ldir + render (nerfactor.py:315-365) vs the reference's own NumPy
    estimator run in the build container (brdf/renderer.py SphereRenderer: light directions,
    cosines, front-lit visibility, light contribution, hemisphere sum) on its sphere scene with a
    Lambertian BRDF."""
    g = np.load(os.path.join(golden_dir, 'ref_sphere_renderer.npz'))
    fg = g['is_fg']
    xyz, normal = torch.tensor(g['xyz'][fg]), torch.tensor(g['normal'][fg])
    albedo = torch.tensor(g['albedo'][fg])
    lvis_fg = torch.ones((int(fg.sum()), g['lxyz'].shape[0] * g['lxyz'].shape[1]),
                         dtype=torch.float64)          # the reference's lvis = fg & front-lit
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 2e-6)):
        om = stage_b.StageB({'light': g['envmap']}, {'brdf': 'microfacet', 'linear2srgb': False},
                            lxyz=g['lxyz'], lareas=g['lareas'], dtype=dtype)
        if dtype == torch.float64:        # StageB casts the lights to fp32 first (shape.py:75-76)
            om.lxyz = torch.tensor(g['lxyz'].reshape(-1, 3))
            om.lareas = torch.tensor(g['lareas'].reshape(-1))
        x, n, a = xyz.to(dtype), normal.to(dtype), albedo.to(dtype)
        l = om.calc_ldir(x)
        cos = torch.einsum('ijk,ik->ij', l, n)
        assert np.abs(cos.double().numpy() - g['lcos'][fg]).max() < (1e-12 if tol < 1e-9 else 1e-6)
        assert np.array_equal((cos > 0).numpy(), g['lvis'][fg].astype(bool)) or tol > 1e-9
        brdf = (a / np.pi)[:, None, :].expand(-1, l.shape[1], -1)
        rgb, _ = om.render(lvis_fg.to(dtype), brdf, l, n)
        want = g['render'][fg]
        err = np.linalg.norm(rgb.double().numpy() - want) / np.linalg.norm(want)
        assert err < tol, (dtype, err)
    assert np.all(g['render'][~fg] == 1.)              # white background of the reference render


# ---- the reference's own model code, run op by op through the TensorFlow shim
# (tests/golden/tfshim + make_golden_tfshim.py, build container) -> ref_tfshim_*.npz
@pytest.mark.parametrize('kind', ['microfacet', 'learned'])
def test_stage_b_oracle_equals_reference_code_via_shim(golden_dir, kind):
    """nerfactor/models/{nerfactor,nerfactor_microfacet,shape,brdf}.py Model.call (test / train /
    vali), compute_loss, the albedo / BRDF edits, OLAT and probe relighting -- the reference's
    files unmodified, TF ops supplied by the shim -- against the oracle on the same inputs."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_b_%s.npz' % kind))
    lh, n = int(g['light_h']), int(g['n_rays'])
    params = synth.make_stage_b_params(int(g['seed_params']), kind, light_hw=(lh, 2 * lh))
    batch = synth.make_stage_b_batch(int(g['seed_batch']), n, 2 * lh * lh)
    om = stage_b.StageB(params, {'brdf': kind}, light_h=lh)
    tol = 1e-6
    op, _, _ = om.call(batch, 'test', relight_lights=list(g['probes']) + om.novel_olat((lh, 2 * lh)))
    for k in ('rgb', 'normal', 'lvis', 'albedo', 'brdf'):
        assert _rel(op[k], g['test_' + k]) < tol, k
    assert _rel(op['rgb_relit'][:, :2], g['test_rgb_probes']) < tol
    assert _rel(op['rgb_relit'][:, 2:], g['test_rgb_olat']) < tol
    op, ogt, olk = om.call(batch, 'train', xyz_noise=g['xyz_noise'])
    for k in ('normal_jitter', 'lvis_jitter', 'albedo_jitter', 'brdf_prop_jitter'):
        assert _rel(olk[k], g['train_' + k]) < tol, k
    wts = {'brdf_smooth_weight': 0.} if kind == 'microfacet' else None   # the two .ini files
    assert _rel(om.compute_loss(op, ogt, weights=wts, **olk), g['train_loss']) < tol
    op, ogt, olk = om.call(batch, 'vali')
    assert _rel(om.compute_loss(op, ogt, weights=wts, **olk), g['vali_loss']) < tol
    kw = {'albedo_override': np.array([0.3, 0.5, 0.7], np.float32)}
    if kind != 'microfacet':
        kw['brdf_z_override'] = np.array([0.01, -0.02, 0.005], np.float32)
    assert _rel(om.call(batch, 'test', **kw)[0]['rgb'], g['edit_rgb']) < tol
    assert _rel(om.call(batch, 'test', albedo_scales=np.array([0.5, 1., 2.], np.float32))[0][
        'rgb'], g['scaled_rgb']) < tol


@pytest.mark.parametrize('kind', ['microfacet', 'learned'])
def test_stage_b_512_lights_oracle_equals_reference_code_via_shim(golden_dir, kind):
    """The reference's native light grid (light_h = 16: 512 lights, nerfactor.ini:49) through the
    shim: forward, two probes and every 16th OLAT render."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_b_%s_L512.npz' % kind))
    lh, n = int(g['light_h']), int(g['n_rays'])
    assert lh == 16
    params = synth.make_stage_b_params(int(g['seed_params']), kind, light_hw=(lh, 2 * lh))
    batch = synth.make_stage_b_batch(int(g['seed_batch']), n, 2 * lh * lh)
    om = stage_b.StageB(params, {'brdf': kind}, light_h=lh)
    op, _, _ = om.call(batch, 'test', relight_lights=list(g['probes']) + om.novel_olat(
        (lh, 2 * lh))[::16])
    for k in ('rgb', 'normal', 'lvis', 'albedo', 'brdf'):
        assert _rel(op[k], g['test_' + k]) < 1e-6, k
    assert _rel(op['rgb_relit'][:, :2], g['test_rgb_probes']) < 1e-6
    assert _rel(op['rgb_relit'][:, 2:], g['test_rgb_olat']) < 1e-6


@pytest.mark.parametrize('kind', ['microfacet', 'learned'])
def test_train_step_gradients_equal_reference_tape_via_shim(golden_dir, kind):
    """trainvali.py:276-285 (forward in train mode, per-ray loss, compute_average_loss,
    tape.gradient over model.trainable_variables) run by the reference's code through the shim
    (GradientTape = torch autograd, tf.custom_gradient honoured: util/math.py safe_acos /
    safe_atan2) vs the oracle's autograd: all 20 Dense layers and the light."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_train_grad_%s.npz' % kind))
    lh, n = int(g['light_h']), int(g['n_rays'])
    params = synth.make_stage_b_params(int(g['seed_params']), kind, light_hw=(lh, 2 * lh))
    batch = synth.make_stage_b_batch(int(g['seed_batch']), n, 2 * lh * lh, fg_frac=1.0)
    tp, leaves = {}, {}
    for k, v in params.items():
        if k == 'light':
            tp[k] = leaves['grad/light'] = torch.tensor(v, requires_grad=True)
            continue
        layers = []
        for li, (w, b) in enumerate(v['layers']):
            wt, bt = torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)
            layers.append((wt, bt))
            leaves['grad/%s/%d/kernel' % (k, li)], leaves['grad/%s/%d/bias' % (k, li)] = wt, bt
        tp[k] = dict(v, layers=layers)
    om = stage_b.StageB(tp, {'brdf': kind, 'shape_mode': 'finetune'}, light_h=lh)
    pred, gt, lk = om.call(batch, 'train', xyz_noise=g['xyz_noise'])
    wts = {'brdf_smooth_weight': 0.} if kind == 'microfacet' else None
    loss = om.compute_loss(pred, gt, weights=wts, **lk)
    assert np.abs(loss.detach().numpy() - g['per_example_loss']).max() < 1e-6
    (loss.sum() / n).backward()
    keys = [k for k in g.files if k.startswith('grad/')]
    assert len(keys) == 41
    for k in keys:
        got, want = leaves[k].grad.numpy(), g[k]
        assert np.abs(got - want).max() <= 1e-6 * max(np.abs(want).max(), 1e-6) + 1e-9, k
    # the frozen BRDF prior gets no gradient in the reference (nerfactor.py:58-60)
    assert not any(k.startswith(('grad/brdf_mlp', 'grad/brdf_out')) for k in keys)


def test_stage_a_oracle_equals_reference_code_via_shim(golden_dir):
    """geometry_from_nerf.{compute_depth_and_normal, compute_light_visibility, eval_sigma_mlp}
    and models/nerf.py call (colour rendering), reference files unmodified via the shim."""
    g = np.load(os.path.join(golden_dir, 'ref_tfshim_stage_a.npz'))
    nerf = synth.make_nerf_params(int(g['seed_nerf']))
    ro, rdn = torch.tensor(g['rayo']), torch.tensor(g['rayd_n'])
    occu, depth, normal = stage_a.compute_depth_and_normal(
        nerf, ro, rdn, 2., 6., n_samples_coarse=-48, n_samples_fine=8)
    assert _rel(occu, g['occu']) < 1e-6 and _rel(depth, g['depth']) < 2e-6
    assert np.abs(normal.numpy() - g['normal']).max() < 5e-5
    surf = ro + rdn * torch.tensor(g['depth'])[:, None]
    lh = int(g['light_h'])
    lxyz, _ = obrdf.gen_light_xyz(lh, 2 * lh)
    lv = stage_a.compute_light_visibility(nerf, surf, torch.tensor(g['normal']),
                                          lxyz.reshape(-1, 3), n_samples_coarse=-48,
                                          n_samples_fine=8)
    assert np.array_equal(np.asarray(lv) != 0, g['lvis_hit'] != 0)       # same front-lit pairs
    assert np.abs(np.asarray(lv) - g['lvis_hit']).max() < 3e-5
    p = torch.tensor(g['sigma_pts'])
    assert _rel(stage_a.eval_sigma_mlp(nerf, p, False), g['sigma_coarse']) < 1e-6
    assert _rel(stage_a.eval_sigma_mlp(nerf, p, True), g['sigma_fine']) < 1e-6
    c, f = stage_a.nerf_render_rays(nerf, ro, torch.tensor(g['rayd']), 2., 6.,
                                    n_samples_coarse=16, n_samples_fine=24)
    for k in ('rgb', 'occu', 'depth'):
        assert _rel(c[k], g['nerf_coarse_' + k]) < 1e-6, k
        assert _rel(f[k], g['nerf_fine_' + k]) < 5e-6, k
    assert _rel(f['disp'], g['nerf_fine_disp']) < 5e-6


def test_oracle_stage_b_goldens_frozen(golden_dir):
    for brdf in ('microfacet', 'learned'):
        g = np.load(os.path.join(golden_dir, 'oracle_stage_b_%s.npz' % brdf))
        lh, lw = int(g['lh']), int(g['lw'])
        lxyz, lareas = obrdf.gen_light_xyz(lh, lw)
        params = synth.make_stage_b_params(int(g['seed_params']), brdf, (lh, lw))
        batch = synth.make_stage_b_batch(int(g['seed_batch']), int(g['n_rays']), lh * lw)
        probes = synth.make_probes(int(g['seed_probes']), 3, (lh, lw))
        m = stage_b.StageB(params, {'brdf': brdf}, lxyz=lxyz, lareas=lareas)
        pred, _, _ = m.call(batch, 'test', relight_lights=[p for p in probes])
        for k in ('rgb', 'normal', 'lvis', 'albedo', 'brdf', 'rgb_relit'):
            assert np.allclose(pred[k].numpy(), g[k], atol=2e-6), (brdf, k)


def test_oracle_fp64_noise_floor():
    """fp32 oracle vs fp64 oracle on the same inputs: the reference's own fp32
    noise floor is far below the 1e-4 rel-L2 acceptance bar."""
    lxyz, lareas = obrdf.gen_light_xyz(2, 8)
    params = synth.make_stage_b_params(7, 'microfacet', (2, 8))
    batch = synth.make_stage_b_batch(11, 64, 16)
    r32 = stage_b.StageB(params, {'brdf': 'microfacet'}, lxyz=lxyz, lareas=lareas
                         ).call(batch, 'test')[0]['rgb']
    r64 = stage_b.StageB(params, {'brdf': 'microfacet'}, lxyz=lxyz, lareas=lareas,
                         dtype=torch.float64).call(batch, 'test')[0]['rgb']
    rel = torch.linalg.norm(r32.double() - r64) / torch.linalg.norm(r64)
    assert rel < 2e-5


def test_oracle_stage_a_goldens_frozen(golden_dir):
    g = np.load(os.path.join(golden_dir, 'oracle_stage_a.npz'))
    nerf = synth.make_nerf_params(int(g['seed_nerf']))
    rayo, rayd = stage_a.gen_rays(synth.look_at_c2w(), synth.CAM_ANGLE_X, 8, 8)
    assert np.array_equal(rayo, g['rayo']) and np.array_equal(rayd, g['rayd'])
    ro = torch.tensor(rayo.reshape(-1, 3))
    rd = stage_a.l2_normalize(torch.tensor(rayd.reshape(-1, 3)), 1)
    sp = stage_a.march_single_pass(nerf, ro, rd, 2., 6., 32)
    assert np.allclose(sp['depth'].numpy(), g['sp_depth'], atol=1e-5)
    assert np.allclose(sp['occu'].numpy(), g['sp_occu'], atol=1e-5)


def test_stage_a_semantics():
    # searchsorted side='right', exclusive cumprod with +1e-6, last delta 1e10
    z = torch.tensor([[1., 2., 3.]])
    sigma = torch.tensor([[0., 1., 0.]])
    w = stage_a.accumulate_sigma(sigma, z, torch.tensor([[0., 0., 1.]]))
    a1 = 1 - np.exp(-1.)
    assert np.allclose(w.numpy(), [[0., a1 * (1 + 1e-6), 0.]], atol=1e-6)
    w2 = stage_a.accumulate_sigma(torch.tensor([[0., 0., 5.]]), z,
                                  torch.tensor([[0., 0., 1.]]))
    assert abs(w2[0, 2].item() - (1 + 1e-6) ** 2) < 1e-5   # alpha=1 via dist=1e10
    # ray index n = y * W + x, no half-pixel offset (datasets/nerf.py:176-193)
    c2w = np.eye(4)
    rayo, rayd = stage_a.gen_rays(c2w, 0.6911, 4, 6)
    fl = .5 * 6 / np.tan(.5 * 0.6911)
    flat = rayd.reshape(-1, 3)
    for (y, x) in ((0, 0), (1, 5), (3, 2)):
        exp = np.array([(x - 3.) / fl, -(y - 2.) / fl, -1.], np.float32)
        assert np.array_equal(flat[y * 6 + x], exp)

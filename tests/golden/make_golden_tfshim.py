"""Runs the UNMODIFIED reference model code (/root/reference) through the TensorFlow shim
(tests/golden/tfshim) on seeded inputs and writes its outputs to tests/golden/ref_tfshim_*.npz.

Build container only (needs /root/reference):

    python tests/golden/make_golden_tfshim.py

The reference decides WHAT is computed -- Model.call / compute_loss of
nerfactor/models/{shape,nerfactor,nerfactor_microfacet}.py, brdf/microfacet/microfacet.py,
geometry_from_nerf.compute_depth_and_normal / compute_light_visibility with models/nerf.py; the
shim only supplies each individual TF op.  The fixtures pin the oracle
(tests/test_oracle_pinning.py) and the CUDA kernels (tests/test_gpu_parity.py) against numbers the
reference's own code produced.
"""
import os
import sys
import tempfile
import warnings
from collections import OrderedDict
from configparser import ConfigParser

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, 'tfshim'), REF, os.path.join(REF, 'nerfactor'), ROOT]
warnings.filterwarnings('ignore')

import refpin  # noqa: E402
refpin.pin()          # the reference's namespace packages (brdf, third_party), not the repo's stubs
import tensorflow as tf  # noqa: E402  (the shim)
import torch  # noqa: E402

assert tf.__version__.endswith('shim')
from nerfactor_b200 import synth  # noqa: E402


def read_ini(name, **override):
    """The reference's own hyper-parameter file, site paths replaced."""
    cfg = ConfigParser()
    with open(os.path.join(REF, 'nerfactor', 'config', name)) as h:
        cfg.read_file(h)
    for k, v in override.items():
        cfg.set('DEFAULT', k, str(v))
    return cfg


def write_ini(cfg, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as h:
        cfg.write(h)


def set_weights(net_dict, params):
    for name, net in net_dict.items():
        if name not in params:
            continue
        layers = params[name]['layers']
        assert len(layers) == len(net.layers), name
        for layer, (w, b) in zip(net.layers, layers):
            layer.set_weights([w, b])


def t32(x):
    return tf.convert_to_tensor(np.asarray(x, np.float32))


def to_np(d):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v)
            for k, v in d.items() if v is not None and not isinstance(v, (str, bytes))}


def build_stage_b(kind, light_h, tmp, seed_params):
    """The reference Model constructed the reference's way: config next to (stub) checkpoints."""
    envdir = os.path.join(tmp, 'envmaps')
    os.makedirs(envdir, exist_ok=True)
    shape_ckpt = os.path.join(tmp, 'shape', 'lr1e-2', 'checkpoints', 'ckpt-2')
    write_ini(read_ini('shape.ini', light_h=light_h), os.path.join(tmp, 'shape', 'lr1e-2.ini'))
    brdf_ckpt = os.path.join(tmp, 'merl', 'lr1e-2', 'checkpoints', 'ckpt-50')
    write_ini(read_ini('brdf.ini', data_root=os.path.join(tmp, 'merl_data')),
              os.path.join(tmp, 'merl', 'lr1e-2.ini'))
    ini = 'nerfactor_microfacet.ini' if kind == 'microfacet' else 'nerfactor.ini'
    cfg = read_ini(ini, light_h=light_h, shape_model_ckpt=shape_ckpt, brdf_model_ckpt=brdf_ckpt,
                   test_envmap_dir=envdir, data_root=tmp, data_nerf_root=tmp, outroot=tmp)
    from nerfactor.models import get_model_class
    Model = get_model_class(cfg.get('DEFAULT', 'model'))
    model = Model(cfg)
    params = synth.make_stage_b_params(seed_params, 'microfacet' if kind == 'microfacet' else
                                       'learned', light_hw=(light_h, 2 * light_h))
    set_weights(model.net, params)
    if kind != 'microfacet':
        set_weights(model.brdf_model.net, params)
    model._light = tf.Variable(t32(params['light']))
    return model, cfg, params


def run_stage_b(kind, light_h, n_rays, seed_params, seed_batch, out_name, slim=False):
    with tempfile.TemporaryDirectory() as tmp:
        model, cfg, params = build_stage_b(kind, light_h, tmp, seed_params)
        L = 2 * light_h * light_h
        batch_np = synth.make_stage_b_batch(seed_batch, n_rays, L)
        probes = synth.make_probes(seed_batch + 1, 2, (light_h, 2 * light_h))
        model.novel_probes = OrderedDict(('p%d' % i, t32(p)) for i, p in enumerate(probes))
        batch = tuple(x if i < 2 else t32(x) for i, x in enumerate(batch_np))
        out = {'kind': kind, 'light_h': light_h, 'n_rays': n_rays, 'seed_params': seed_params,
               'seed_batch': seed_batch, 'probes': probes}
        # ---- test mode (jitter drawn by the reference, irrelevant to pred)
        pred, gt, lk, _ = model.call(batch, mode='test', relight_olat=True, relight_probes=True)
        for k, v in to_np(pred).items():
            out['test_' + k] = v
        if slim:            # the 512-light case: forward + relighting only (file size)
            out['test_rgb_olat'] = out['test_rgb_olat'][:, ::16]        # every 16th OLAT
            np.savez_compressed(os.path.join(HERE, out_name), **out)
            print(out_name, 'slim', out['test_rgb_olat'].shape)
            return out
        # ---- train mode with the jitter noise recorded (nerfactor.py:198-201) + compute_loss
        tf.random.set_seed(777)
        pred, gt, lk, _ = model.call(batch, mode='train')
        tf.random.set_seed(777)
        n_fg = int((batch_np[5][:, 0] > 0).sum())
        out['xyz_noise'] = tf.random.normal(
            (n_fg, 3), stddev=cfg.getfloat('DEFAULT', 'xyz_jitter_std')).numpy()
        for k, v in to_np(pred).items():
            out['train_' + k] = v
        for k, v in to_np(lk).items():
            out['train_' + k] = v
        lk['keep_batch'] = True
        out['train_loss'] = model.compute_loss(pred, gt, **lk).numpy()
        pred_v, gt_v, lk_v, _ = model.call(batch, mode='vali')
        out['vali_loss'] = model.compute_loss(pred_v, gt_v, **lk_v).numpy()
        # ---- edits (test.py:91-132, 168-186)
        alb_o = np.array([0.3, 0.5, 0.7], np.float32)
        kw = {'albedo_override': t32(alb_o), 'albedo_scales': None}
        if kind != 'microfacet':
            kw['brdf_z_override'] = t32([0.01, -0.02, 0.005])
        pred_e, _, _, _ = model.call(batch, mode='test', **kw)
        out['edit_rgb'] = pred_e['rgb'].numpy()
        pred_s, _, _, _ = model.call(batch, mode='test', albedo_scales=t32([0.5, 1., 2.]))
        out['scaled_rgb'] = pred_s['rgb'].numpy()
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, {k: getattr(v, 'shape', v) for k, v in out.items() if k.startswith('test_')})
    return out


def run_train_gradients(kind, light_h, n_rays, seed_params, seed_batch, out_name):
    """One `train_step` of nerfactor/trainvali.py:276-285 up to the gradients: forward in train
    mode, per-ray loss, tf.nn.compute_average_loss, tape.gradient over model.trainable_variables
    (shape_mode finetune: all four shape MLPs, albedo, BRDF-z / roughness MLPs and the light)."""
    tf.shim_set_training(True)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            model, cfg, params = build_stage_b(kind, light_h, tmp, seed_params)
            model.register_trainable()
            L = 2 * light_h * light_h
            batch_np = synth.make_stage_b_batch(seed_batch, n_rays, L, fg_frac=1.0)
            batch = tuple(x if i < 2 else t32(x) for i, x in enumerate(batch_np))
            tf.random.set_seed(4242)
            with tf.GradientTape() as tape:                       # trainvali.py:277-283
                pred, gt, loss_kwargs, _ = model(batch, mode='train')
                loss_kwargs['keep_batch'] = True
                per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
                weighted_loss = tf.nn.compute_average_loss(
                    per_example_loss, global_batch_size=n_rays)
            variables = model.trainable_variables
            grads = tape.gradient(weighted_loss, variables)       # trainvali.py:284
            tf.random.set_seed(4242)
            noise = tf.random.normal((n_rays, 3), stddev=cfg.getfloat('DEFAULT', 'xyz_jitter_std'))
            out = {'kind': kind, 'light_h': light_h, 'n_rays': n_rays, 'seed_params': seed_params,
                   'seed_batch': seed_batch, 'xyz_noise': noise.numpy(),
                   'per_example_loss': per_example_loss.detach().numpy(),
                   'weighted_loss': float(weighted_loss)}
            # name every gradient by the owner of its variable
            owner = {}
            for net_name, net in model.net.items():
                for li, layer in enumerate(net.layers):
                    owner[id(layer.kernel)] = 'grad/%s/%d/kernel' % (net_name, li)
                    owner[id(layer.bias)] = 'grad/%s/%d/bias' % (net_name, li)
            owner[id(model._light)] = 'grad/light'
            for v, g_ in zip(variables, grads):
                assert g_ is not None, owner[id(v)]
                out[owner[id(v)]] = g_.numpy()
    finally:
        tf.shim_set_training(False)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, len([k for k in out if k.startswith('grad/')]), 'gradient tensors, loss',
          out['weighted_loss'])
    return out


def run_nerf_train_gradients(seed_nerf, n_rays, out_name):
    """The NeRF's own train step (models/nerf.py call + compute_loss under trainvali.py:276-285)
    with stratified perturbation, importance sampling with random u and density noise ON; the
    four random draws are recorded in call order (gen_z uniform, coarse density normal,
    gen_z_fine uniform, fine density normal)."""
    from nerfactor.models.nerf import Model as NerfModel
    n_c, n_f, noise_std = 8, 6, 0.5
    tf.shim_set_training(True)
    try:
        cfg = read_ini('nerf.ini', n_samples_coarse=n_c, n_samples_fine=n_f, perturb=True,
                       noise_std=noise_std, data_root='/tmp', outroot='/tmp')
        model = NerfModel(cfg)
        params = synth.make_nerf_params(seed_nerf)
        set_weights(model.net, params)
        model.register_trainable()
        rng = np.random.default_rng(21)
        c2w = synth.look_at_c2w()
        from oracle import stage_a
        rayo, rayd = stage_a.gen_rays(c2w, synth.CAM_ANGLE_X, 8, 8)
        sel = rng.choice(64, n_rays, replace=False)
        rayo, rayd = rayo.reshape(-1, 3)[sel], rayd.reshape(-1, 3)[sel]
        rgb = rng.uniform(0, 1, (n_rays, 3)).astype(np.float32)
        batch = (None, None, t32(rayo), t32(rayd), t32(rgb))
        tf.random.set_seed(31337)
        recorded = {}
        orig_gen_z_fine = NerfModel.gen_z_fine          # record what the reference sampled

        def recording_gen_z_fine(*a, **k):
            recorded['z_all'] = orig_gen_z_fine(*a, **k)
            return recorded['z_all']
        NerfModel.gen_z_fine = staticmethod(recording_gen_z_fine)
        with tf.GradientTape() as tape:
            pred, gt, loss_kwargs, _ = model(batch, mode='train')
            loss_kwargs['keep_batch'] = True
            per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
            weighted_loss = tf.nn.compute_average_loss(per_example_loss,
                                                       global_batch_size=n_rays)
        NerfModel.gen_z_fine = staticmethod(orig_gen_z_fine)
        variables = model.trainable_variables
        grads = tape.gradient(weighted_loss, variables)
        tf.random.set_seed(31337)                       # replay the draws in call order
        u1 = tf.random.uniform((n_rays, n_c))
        g1 = tf.random.normal((n_rays, n_c)) * noise_std
        u2 = tf.random.uniform((n_rays, n_f))
        g2 = tf.random.normal((n_rays, n_c + n_f)) * noise_std
        out = {'seed_nerf': seed_nerf, 'n_c': n_c, 'n_f': n_f, 'noise_std': noise_std,
               'rayo': rayo, 'rayd': rayd, 'rgb': rgb, 'perturb_u': u1.numpy(),
               'fine_u': u2.numpy(), 'noise_coarse': g1.numpy(), 'noise_fine': g2.numpy(),
               'z_all': recorded['z_all'].detach().numpy(),
               'pred_coarse': pred['coarse'].numpy(), 'pred_fine': pred['fine'].numpy(),
               'per_example_loss': per_example_loss.numpy()}
        owner = {}
        for net_name, net in model.net.items():
            for li, layer in enumerate(net.layers):
                owner[id(layer.kernel)] = 'grad/%s/%d/kernel' % (net_name, li)
                owner[id(layer.bias)] = 'grad/%s/%d/bias' % (net_name, li)
        for v, g_ in zip(variables, grads):
            out[owner[id(v)]] = g_.numpy().astype(np.float16 if g_.numel() > 40000 else np.float32)
    finally:
        tf.shim_set_training(False)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, len([k for k in out if k.startswith('grad/')]), 'gradient tensors')
    return out


def run_brdf_train_gradients(out_name):
    """The BRDF prior's train step: nerfactor/models/brdf.py call + compute_loss (log-space L2,
    reciprocity term) under trainvali.py:276-285; gradients of the MLP and of the latent codes."""
    from nerfactor.models.brdf import Model as BrdfModel
    tf.shim_set_training(True)
    try:
        with tempfile.TemporaryDirectory() as tmp:
            names = synth.write_merl_npz(tmp, n_rows=64)
            cfg = read_ini('brdf.ini', data_root=tmp, outroot=tmp)
            model = BrdfModel(cfg)
            assert model.brdf_names == sorted(names)
            params = synth.make_stage_b_params(5, 'learned')
            set_weights(model.net, params)
            z0 = (0.01 * np.random.default_rng(8).standard_normal((len(names), 3))).astype(np.float32)
            model.latent_code.z = t32(z0)
            model.register_trainable()
            d = dict(np.load(os.path.join(tmp, 'train_%s.npz' % model.brdf_names[1])))
            n = d['rusink'].shape[0]
            batch = (None, tf.convert_to_tensor(np.full((n,), int(d['i']), np.int32)), None, None,
                     None, t32(d['rusink']), t32(d['refl']))
            with tf.GradientTape() as tape:
                pred, gt, loss_kwargs, _ = model(batch, mode='train')
                loss_kwargs['keep_batch'] = True
                per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
                weighted_loss = tf.nn.compute_average_loss(per_example_loss, global_batch_size=n)
            variables = model.trainable_variables
            grads = tape.gradient(weighted_loss, variables)
            out = {'names': np.array(model.brdf_names), 'i': int(d['i']), 'z0': z0,
                   'rusink': d['rusink'], 'refl': d['refl'], 'pred_brdf': pred['brdf'].numpy(),
                   'pred_brdf_reci': pred['brdf_reci'].numpy(),
                   'per_example_loss': per_example_loss.numpy()}
            owner = {id(model.latent_code._z): 'grad/z'}
            for net_name, net in model.net.items():
                for li, layer in enumerate(net.layers):
                    owner[id(layer.kernel)] = 'grad/%s/%d/kernel' % (net_name, li)
                    owner[id(layer.bias)] = 'grad/%s/%d/bias' % (net_name, li)
            for v, g_ in zip(variables, grads):
                out[owner[id(v)]] = g_.numpy()
            assert 'grad/z' in out
    finally:
        tf.shim_set_training(False)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, len([k for k in out if k.startswith('grad/')]), 'gradient tensors')
    return out


def run_shape(light_h, n_rays, seed_params, seed_batch, out_name):
    """nerfactor/models/shape.py Model.call (train, recorded jitter) + compute_loss."""
    from nerfactor.models.shape import Model
    # shape.ini ships without smoothness weights (fallback 0, shape.py:37-40); set them so the
    # jittered branch is exercised too
    cfg = read_ini('shape.ini', light_h=light_h, data_root='/tmp', data_nerf_root='/tmp',
                   outroot='/tmp', normal_smooth_weight=0.01, lvis_smooth_weight=0.5)
    model = Model(cfg)
    params = synth.make_stage_b_params(seed_params, 'learned', light_hw=(light_h, 2 * light_h))
    set_weights(model.net, params)
    L = 2 * light_h * light_h
    batch_np = synth.make_stage_b_batch(seed_batch, n_rays, L)
    batch = tuple(x if i < 2 else t32(x) for i, x in enumerate(batch_np))
    tf.random.set_seed(99)
    pred, gt, lk, _ = model.call(batch, mode='train')
    tf.random.set_seed(99)
    noise = tf.random.normal((n_rays, 3), stddev=cfg.getfloat('DEFAULT', 'xyz_jitter_std'))
    out = {'light_h': light_h, 'n_rays': n_rays, 'seed_params': seed_params,
           'seed_batch': seed_batch, 'xyz_noise': noise.numpy()}
    for k, v in to_np(pred).items():
        out['pred_' + k] = v
    for k, v in to_np(lk).items():
        out[k] = v
    out['loss'] = model.compute_loss(pred, gt, **lk).numpy()
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, sorted(out))
    return out


def run_stage_a(seed_nerf, hw, light_h, out_name):
    """geometry_from_nerf.compute_depth_and_normal / compute_light_visibility / eval_sigma_mlp and
    the NeRF colour rendering (models/nerf.py call) of the reference, on a random-init NeRF."""
    from nerfactor import geometry_from_nerf as gfn
    from nerfactor.models.nerf import Model as NerfModel
    from oracle import stage_a                       # only for its NumPy ray generator
    if not gfn.FLAGS.is_parsed():
        gfn.FLAGS(['make_golden_tfshim', '--light_h=%d' % light_h])
    cfg = read_ini('nerf.ini', n_samples_coarse=-48, n_samples_fine=8, data_root='/tmp',
                   outroot='/tmp')                   # geometry_from_nerf adds 64 to both
    model = NerfModel(cfg)
    params = synth.make_nerf_params(seed_nerf)
    set_weights(model.net, params)
    h, w = hw
    rayo, rayd = stage_a.gen_rays(synth.look_at_c2w(), synth.CAM_ANGLE_X, h, w)
    rayo, rayd = rayo.reshape(-1, 3), rayd.reshape(-1, 3)
    rayd_n = tf.linalg.l2_normalize(t32(rayd), axis=1)                  # gfn.py:100
    out = {'seed_nerf': seed_nerf, 'h': h, 'w': w, 'light_h': light_h, 'rayo': rayo,
           'rayd': rayd, 'rayd_n': rayd_n.numpy(), 'n_samples_coarse': 16, 'n_samples_fine': 72}
    occu, depth, normal = [x.detach() for x in gfn.compute_depth_and_normal(
        model, t32(rayo), rayd_n, cfg)]       # values only (the tape's graph is torch autograd here)
    out.update(occu=occu.numpy(), depth=depth.numpy(), normal=normal.numpy())
    surf = t32(rayo) + rayd_n * depth[:, None]                          # gfn.py:134
    out['lvis_hit'] = gfn.compute_light_visibility(model, surf, normal, cfg)
    pts = t32(np.random.default_rng(5).uniform(-1.5, 1.5, size=(257, 3)))
    out['sigma_pts'] = pts.numpy()
    out['sigma_coarse'] = gfn.eval_sigma_mlp(model, pts, use_fine=False).numpy()
    out['sigma_fine'] = gfn.eval_sigma_mlp(model, pts, use_fine=True).numpy()
    # NeRF colour rendering with the config's own sample counts (16 + 24 here)
    cfg2 = read_ini('nerf.ini', n_samples_coarse=16, n_samples_fine=24, data_root='/tmp',
                    outroot='/tmp')
    model2 = NerfModel(cfg2)
    set_weights(model2.net, params)
    batch = (np.array([b'v'] * (h * w)), np.tile(np.array([[h, w]], np.int32), (h * w, 1)),
             t32(rayo), t32(rayd), t32(np.zeros((h * w, 3))))
    pred, _, _, to_vis = model2.call(batch, mode='test')
    for k in ('coarse_rgb', 'coarse_occu', 'coarse_depth', 'fine_rgb', 'fine_occu', 'fine_depth',
              'fine_disp'):
        out['nerf_' + k] = to_vis[k].numpy()
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, 'occu range', float(out['occu'].min()), float(out['occu'].max()),
          'front-lit pairs', int((out['lvis_hit'] != 0).sum()))
    return out


if __name__ == '__main__':
    if 'l512' in sys.argv[1:]:
        run_stage_b('microfacet', 16, 40, 21, 22, 'ref_tfshim_stage_b_microfacet_L512.npz', slim=True)
        run_stage_b('learned', 16, 40, 21, 22, 'ref_tfshim_stage_b_learned_L512.npz', slim=True)
        sys.exit(0)
    if 'brdfgrad' in sys.argv[1:]:
        run_brdf_train_gradients('ref_tfshim_brdf_train_grad.npz')
        sys.exit(0)
    if 'nerfgrad' in sys.argv[1:]:
        run_nerf_train_gradients(3, 12, 'ref_tfshim_nerf_train_grad.npz')
        sys.exit(0)
    if 'shape' in sys.argv[1:]:
        run_shape(2, 40, 3, 9, 'ref_tfshim_shape.npz')
        sys.exit(0)
    if 'grad' in sys.argv[1:]:
        run_train_gradients('microfacet', 2, 48, 7, 11, 'ref_tfshim_train_grad_microfacet.npz')
        run_train_gradients('learned', 2, 48, 7, 11, 'ref_tfshim_train_grad_learned.npz')
        sys.exit(0)
    run_stage_a(3, (6, 6), 2, 'ref_tfshim_stage_a.npz')
    run_shape(2, 40, 3, 9, 'ref_tfshim_shape.npz')
    run_nerf_train_gradients(3, 12, 'ref_tfshim_nerf_train_grad.npz')
    run_brdf_train_gradients('ref_tfshim_brdf_train_grad.npz')
    run_stage_b('microfacet', 4, 80, 7, 11, 'ref_tfshim_stage_b_microfacet.npz')
    run_stage_b('learned', 4, 80, 7, 11, 'ref_tfshim_stage_b_learned.npz')
    run_stage_b('microfacet', 16, 40, 21, 22, 'ref_tfshim_stage_b_microfacet_L512.npz', slim=True)
    run_stage_b('learned', 16, 40, 21, 22, 'ref_tfshim_stage_b_learned_L512.npz', slim=True)
    run_train_gradients('microfacet', 2, 48, 7, 11, 'ref_tfshim_train_grad_microfacet.npz')
    run_train_gradients('learned', 2, 48, 7, 11, 'ref_tfshim_train_grad_learned.npz')

"""Seeded synthetic scenes, weights and batches.

There is no network for datasets or checkpoints, so every benchmark and parity
test runs on random-init networks of the reference architecture and synthetic
geometry of the reference batch layout
(nerfactor/datasets/nerf_shape.py:72-95: id_, hw, rayo, rayd, rgb, alpha, xyz,
normal, lvis).  Pure NumPy; no oracle, no CUDA.
"""
import math

import numpy as np

CAM_ANGLE_X = 0.6911  # Blender-synthetic-like horizontal FoV


def look_at_c2w(radius=4.0, azimuth_deg=30.0, elevation_deg=30.0):
    """Camera-to-world 4x4 (OpenGL convention: camera looks down -z, +y up),
    the layout of metadata['cam_transform_mat'] (datasets/nerf.py:146-148)."""
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    eye = radius * np.array(
        [math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0., 0., 1.])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    true_up = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, true_up, -fwd, eye
    return c2w


def glorot_uniform(rng, fan_in, fan_out):
    """Keras Dense default (nerfactor/networks/mlp.py:34)."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def init_mlp(rng, in_dim, widths, act, skip_at=None, bias_std=0.0):
    """Parameter dict {'layers': [(W[in,out], b[out])...], 'act', 'skip_at'} with
    the input sizes Keras would infer for mlp.Network (mlp.py:39-50)."""
    layers, d = [], in_dim
    for i, w in enumerate(widths):
        W = glorot_uniform(rng, d, w)
        b = (rng.standard_normal(w) * bias_std).astype(np.float32)
        layers.append((W, b))
        d = w + in_dim if (skip_at is not None and i in skip_at) else w
    return {'layers': layers, 'act': list(act) if act else [None] * len(widths),
            'skip_at': skip_at}


def embed_dims(n_freqs, in_dims=3):
    return in_dims * (1 + 2 * n_freqs)


def make_stage_b_params(seed=0, brdf='microfacet', light_hw=(16, 32), width=128,
                        depth=4, skip_at=2, n_freqs_xyz=10, n_freqs_ldir=4,
                        n_freqs_rusink=2, z_dim=3, bias_std=0.05, xyz_freq_decay=0.0):
    """Random-init NeRFactor networks (nerfactor.py:128-167, shape.py:79-94,
    brdf.py:57-66, nerfactor_microfacet.py:108-114) + light (nerfactor.py:367-375).

    `xyz_freq_decay` d > 0 scales the weights that read octave f of the xyz encoding by
    2^(-d f) (layer 0 and the skip layer): the spectrum of a *trained* network, whose outputs
    vary smoothly with the surface point.  With d = 0 (Keras init) the outputs change by O(1)
    when xyz moves by 2^-9, which makes any comparison across two Stage-A implementations
    meaningless; end-to-end tests use d = 1."""
    rng = np.random.default_rng(seed)
    dx, dl, dr = embed_dims(n_freqs_xyz), embed_dims(n_freqs_ldir), \
        embed_dims(n_freqs_rusink)
    trunk = lambda d_in: init_mlp(
        rng, d_in, [width] * depth, ['relu'] * depth, [skip_at], bias_std)
    p = {}
    p['normal_mlp'] = trunk(dx)
    p['normal_out'] = init_mlp(rng, width, [3], [None], None, bias_std)
    p['lvis_mlp'] = trunk(dx + dl)
    p['lvis_out'] = init_mlp(rng, width, [1], ['sigmoid'], None, bias_std)
    p['albedo_mlp'] = trunk(dx)
    p['albedo_out'] = init_mlp(rng, width, [3], ['sigmoid'], None, bias_std)
    p['brdf_z_mlp'] = trunk(dx)
    if brdf == 'microfacet':
        p['brdf_z_out'] = init_mlp(rng, width, [1], ['sigmoid'], None, bias_std)
    else:
        p['brdf_z_out'] = init_mlp(rng, width, [z_dim], [None], None, bias_std)
        p['brdf_mlp'] = trunk(z_dim + dr)
        p['brdf_out'] = init_mlp(rng, width, [1], ['softplus'], None, bias_std)
    p['light'] = rng.uniform(0., 1., size=light_hw + (3,)).astype(np.float32)
    if xyz_freq_decay > 0:
        for name in ('normal_mlp', 'lvis_mlp', 'albedo_mlp', 'brdf_z_mlp'):
            layers = p[name]['layers']
            for li, row0 in ((0, 0), (skip_at + 1, width)):
                W, b = layers[li]
                W = W.copy()
                for f in range(n_freqs_xyz):
                    W[row0 + 3 + 6 * f:row0 + 9 + 6 * f] *= 2.0 ** (-xyz_freq_decay * f)
                layers[li] = (W, b)
    return p


def make_nerf_params(seed=0, width=256, enc_depth=8, n_freqs_xyz=10,
                     bias_std=0.05, sigma_gain=40.0, sigma_bias=8.0, n_freqs_view=4):
    """Random-init NeRF sigma networks (models/nerf.py:53-71: enc = 8x256 ReLU
    with the input re-concatenated after layer enc_depth//2, sigma_out =
    Dense(1)); coarse and fine copies (nerf.py:42-47).  A glorot-init field is a
    thin uniform fog, so the output layer is rescaled (sigma_gain) and shifted
    (sigma_bias) to give a scene-like field: ~1/3 of space occupied, per-step
    opacities of a few tenths."""
    rng = np.random.default_rng(seed)
    dx = embed_dims(n_freqs_xyz)
    p = {}
    for pref in ('coarse_', 'fine_'):
        p[pref + 'enc'] = init_mlp(
            rng, dx, [width] * enc_depth, ['relu'] * enc_depth,
            [enc_depth // 2], bias_std)
        so = init_mlp(rng, width, [1], [None], None, 0.0)
        so['layers'][0] = ((so['layers'][0][0] * sigma_gain).astype(np.float32),
                           np.full((1,), sigma_bias, np.float32))
        p[pref + 'sigma_out'] = so
    # colour branch (nerf.py:66-70); its own generator so the sigma networks above -- and every
    # golden fixture made from them -- do not depend on it
    rng2 = np.random.default_rng(100003 + seed)
    dv = embed_dims(n_freqs_view)
    for pref in ('coarse_', 'fine_'):
        p[pref + 'bottleneck'] = init_mlp(rng2, width, [width], [None], None, bias_std)
        p[pref + 'rgb_out'] = init_mlp(rng2, width + dv, [width // 2, 3], ['relu', None], None,
                                       bias_std)
    return p


def make_stage_b_batch(seed, n_rays, n_lights, fg_frac=0.75, cam_loc=None):
    """The 9-tuple `Model.call` consumes (SURVEY 8d 'Stage B inputs')."""
    rng = np.random.default_rng(seed)
    if cam_loc is None:
        cam_loc = look_at_c2w()[:3, 3]
    xyz = rng.uniform(-1., 1., size=(n_rays, 3))
    r = rng.uniform(0.5, 1.5, size=(n_rays, 1))
    xyz = xyz / np.maximum(np.linalg.norm(xyz, axis=1, keepdims=True), 1e-3) * r
    normal = xyz / np.linalg.norm(xyz, axis=1, keepdims=True) + \
        0.1 * rng.standard_normal((n_rays, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    alpha = (rng.uniform(size=(n_rays, 1)) < fg_frac).astype(np.float32)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    id_ = np.array([b'synth_000'] * n_rays)
    hw = np.tile(np.array([[n_rays, 1]], np.int32), (n_rays, 1))
    rayo = np.tile(f32(cam_loc)[None, :], (n_rays, 1))
    rayd = f32(xyz) - rayo
    rgb = rng.uniform(0., 1., size=(n_rays, 3))
    lvis = rng.uniform(0., 1., size=(n_rays, n_lights))
    return (id_, hw, rayo, f32(rayd), f32(rgb), f32(alpha), f32(xyz), f32(normal),
            f32(lvis))


def make_probes(seed, n_probes, light_hw=(16, 32)):
    """Synthetic HDR light probes: exp(N(0, 1.5^2)) clipped to [0, 200]."""
    rng = np.random.default_rng(seed)
    x = np.exp(rng.normal(0., 1.5, size=(n_probes,) + tuple(light_hw) + (3,)))
    return np.clip(x, 0., 200.).astype(np.float32)


def light_index_map(envmap_hw, light_hw):
    """Nearest-pixel map from each of the light_hw[0]*light_hw[1] directions to a
    pixel of a (smaller or equal) env-map (SURVEY 8d caveat on L): identity when
    the two grids agree."""
    eh, ew = envmap_hw
    lh, lw = light_hw
    ii = (np.arange(lh) * eh) // lh
    jj = (np.arange(lw) * ew) // lw
    return (ii[:, None] * ew + jj[None, :]).reshape(-1).astype(np.int32)


def write_scene(root, imh=16, imw=16, n_train=2, n_val=1, n_test=2, seed=0, nerf_root=None,
                n_lights=None, envmap_dir=None, n_probes=2, light_hw=(16, 32)):
    """Writes a tiny synthetic scene in the reference's on-disk layout (data_gen output,
    datasets/nerf.py:64-90): `<root>/{train,val,test}_NNN/metadata.json` (+ `rgba.png` for
    train / val: a shaded sphere on a transparent background), optionally the Stage-A buffers
    `<nerf_root>/<view>/{alpha.png,xyz.npy,normal.npy,lvis.npy}` (analytic sphere of radius 1)
    and `n_probes` Radiance .hdr light probes in `envmap_dir`.  Returns the list of view ids."""
    import json
    import os
    from PIL import Image
    rng = np.random.default_rng(seed)
    ids = []
    k = 0
    for mode, n in (('train', n_train), ('val', n_val), ('test', n_test)):
        for i in range(n):
            id_ = '%s_%03d' % (mode, i)
            ids.append(id_)
            c2w = look_at_c2w(4.0, 30.0 + 40.0 * k, 20.0 + 5.0 * k)
            k += 1
            d = os.path.join(root, id_)
            os.makedirs(d, exist_ok=True)
            meta = {'id': id_, 'imh': imh, 'imw': imw, 'cam_angle_x': CAM_ANGLE_X,
                    'cam_transform_mat': ','.join('%.17g' % v for v in c2w.reshape(-1))}
            with open(os.path.join(d, 'metadata.json'), 'w') as f:
                json.dump(meta, f)
            # analytic unit sphere seen from this camera
            fl = .5 * imw / np.tan(.5 * CAM_ANGLE_X)
            xs, ys = np.meshgrid(np.arange(imw, dtype=float), np.arange(imh, dtype=float))
            dl = np.stack(((xs - .5 * imw) / fl, -(ys - .5 * imh) / fl, -np.ones_like(xs)), -1)
            dw = np.sum(dl[:, :, None, :] * c2w[:3, :3], -1)
            dn = dw / np.linalg.norm(dw, axis=-1, keepdims=True)
            o = c2w[:3, 3]
            b = dn @ o
            disc = b * b - (o @ o - 1.0)
            hit = disc > 0
            t = -b - np.sqrt(np.where(hit, disc, 0.))
            p = o[None, None, :] + dn * t[..., None]
            nrm = np.where(hit[..., None], p, np.array([0., 1., 0.]))
            alpha = hit.astype(np.float32)
            if mode != 'test':
                shade = np.clip(nrm @ np.array([0.3, 0.5, 0.8]), 0.05, 1.)[..., None]
                rgb = shade * np.array([0.8, 0.6, 0.4])
                rgba = np.concatenate([rgb * alpha[..., None], alpha[..., None]], -1)
                Image.fromarray((np.clip(rgba, 0, 1) * 255).astype(np.uint8), 'RGBA').save(
                    os.path.join(d, 'rgba.png'))
            if nerf_root is not None:
                bd = os.path.join(nerf_root, id_)
                os.makedirs(bd, exist_ok=True)
                L = n_lights or light_hw[0] * light_hw[1]
                Image.fromarray((alpha * 255).astype(np.uint8)).save(os.path.join(bd, 'alpha.png'))
                np.save(os.path.join(bd, 'xyz.npy'), (p * alpha[..., None]).astype(np.float32))
                np.save(os.path.join(bd, 'normal.npy'), nrm.astype(np.float32))
                np.save(os.path.join(bd, 'lvis.npy'),
                        (rng.uniform(size=(imh, imw, L)) * alpha[..., None]).astype(np.float32))
    if envmap_dir is not None:
        import cv2
        os.makedirs(envmap_dir, exist_ok=True)
        probes = make_probes(seed + 1, n_probes, (4 * light_hw[0], 4 * light_hw[1]))
        for i, pr in enumerate(probes):
            cv2.imwrite(os.path.join(envmap_dir, 'probe%d.hdr' % i),
                        np.ascontiguousarray(pr[:, :, ::-1]))
    return ids


def write_merl_npz(root, names=('alum-bronze', 'blue-rubber', 'gold-paint'), n_rows=400, seed=0):
    """Synthetic stand-ins for the BRDF prior's training files (data_gen/merl/make_dataset.py
    output, read by datasets/brdf_merl.py): `train_<name>.npz` / `vali_<name>.npz` with Rusinkiewicz
    coordinates (phi_d, theta_h, theta_d) and positive reflectance, and one shared `test.npz`."""
    import os
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)

    def coords(n):
        return np.stack((rng.uniform(0, np.pi, n), rng.uniform(0, np.pi / 2, n),
                         rng.uniform(0, np.pi / 2, n)), 1).astype(np.float32)
    for i, name in enumerate(names):
        for split in ('train', 'vali'):
            r = coords(n_rows)
            refl = np.exp(rng.normal(-1. + 0.3 * i, 0.8, (n_rows, 1))).astype(np.float32)
            np.savez(os.path.join(root, '%s_%s.npz' % (split, name)), name=name, i=i, envmap_h=16,
                     ims=128, spp=1, rusink=r, refl=refl)
    np.savez(os.path.join(root, 'test.npz'), envmap_h=16, ims=128, spp=1, rusink=coords(n_rows))
    return list(names)


def make_blob_nerf_params(seed=0, radius=1.0, sharpness=6.0, gain=37.3, noise=0.35,
                          width=256, enc_depth=8, n_freqs_xyz=10):
    """The analytic 'sphere-like' density field of SURVEY.md 8d inside the reference
    architecture (models/nerf.py:53-71), for tests where depth has to be well-conditioned:

        sigma(x) = relu(gain * carry^7 * relu(sharpness * (sum_i cos(x_i) - c)) + noise-net(x) - 1)

    `sum_i cos(x_i)` is three columns of the positional encoding (octave 0), so its level set
    through (radius, 0, 0) -- a rounded sphere -- is computed by unit 0 of layer 0; units 0 of
    the other layers carry it to the head with a non-dyadic weight (so 16-bit operand rounding is
    exercised at every layer), and the remaining 255 units are a random-init network whose
    output perturbs sigma by about +-`noise` (empty space stays empty: head bias -1).  Coarse and
    fine networks are the same field with different random parts."""
    p = make_nerf_params(seed, width, enc_depth, n_freqs_xyz, sigma_gain=1.0, sigma_bias=0.0)
    c = 2.0 + math.cos(radius)
    carry = 0.973
    for pref in ('coarse_', 'fine_'):
        layers = p[pref + 'enc']['layers']
        for li, (W, b) in enumerate(layers):
            W, b = W.copy(), b.copy()
            W[:, 0] = 0.
            if li == 0:
                W[6:9, 0] = sharpness                    # cos(2^0 x), cos(2^0 y), cos(2^0 z)
                b[0] = -sharpness * c
            else:
                W[0, 0] = carry
                b[0] = 0.
            layers[li] = (W.astype(np.float32), b.astype(np.float32))
        w_out, _ = p[pref + 'sigma_out']['layers'][0]
        w_out = w_out.copy() * noise * 4.0
        w_out[0, 0] = gain / carry ** (enc_depth - 1)
        p[pref + 'sigma_out']['layers'][0] = (w_out.astype(np.float32),
                                              np.full((1,), -1.0, np.float32))
    return p

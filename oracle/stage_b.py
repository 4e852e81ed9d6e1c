"""Oracle restatement of Stage B: the NeRFactor render-and-relight forward
(nerfactor/models/{nerfactor,nerfactor_microfacet,shape}.py). Test infrastructure.

`StageB` mirrors `Model.call` and its helpers on PyTorch-CPU tensors.  Random
quantities of the reference (xyz jitter, `nerfactor.py:199`) are explicit inputs.
"""
import math

import numpy as np
import torch

from . import networks as nets
from . import brdf as brdfmod
from .tfops import safe_l2_normalize


DEFAULT_CONFIG = {
    # nerfactor/config/nerfactor.ini (values) / code fallbacks noted in SURVEY 8a
    'xyz_scale': 1.0,            # shape.py:47-48 fallback
    'albedo_slope': 0.77,        # nerfactor.ini:64
    'albedo_bias': 0.03,         # nerfactor.ini:65
    'n_freqs_xyz': 10,           # nerfactor.ini:87
    'n_freqs_ldir': 4,           # nerfactor.ini:88
    'n_freqs_rusink': 2,         # brdf.ini:47
    'z_dim': 3,                  # brdf.ini:50
    'normalize_brdf_z': False,   # brdf.ini:53
    'learned_brdf_scale': 1.0,   # nerfactor.ini:71
    'linear2srgb': True,         # nerfactor.ini:76
    'fresnel_f0': 0.04,          # nerfactor_microfacet.ini:69
    'shape_mode': 'finetune',    # nerfactor.ini:56
    'brdf': 'learned',           # 'learned' (nerfactor.py) | 'microfacet'
    'white_bg': True,            # nerfactor.ini:50
    'xyz_jitter_std': 0.01,      # nerfactor.ini:53
    'mlp_chunk': 65536,          # nerfactor.ini:81
    'olat_inten': 200.0,         # nerfactor.py:68
    'ambient_inten': 0.0,        # nerfactor.py:69-70
}


def linear2srgb(x):
    """nerfactor/util/img.py:140-163: pow on all elements, then select."""
    x = torch.clamp(x, 0., 1.)
    lin = x * 12.92
    # forward: identical to pow(x, 1/2.4) wherever that branch is selected; the floor only
    # keeps the UNSELECTED branch's derivative finite at x = 0 (TF's pow gradient uses
    # mul_no_nan there), so autograd gives the reference's gradient instead of 0 * inf
    nonlin = 1.055 * torch.pow(torch.clamp(x, min=1e-30), 1 / 2.4) - (1.055 - 1)
    return torch.where(x <= 0.0031308, lin, nonlin)


def alpha_blend(t1, alpha, t2=None):
    """nerfactor/util/img.py:76-95."""
    if t2 is None:
        t2 = torch.zeros_like(t1)
    return t1 * alpha + t2 * (1. - alpha)


def one_hot_img(h, w, c, i, j, dtype=torch.float32):
    """nerfactor/util/tensor.py:57-64."""
    img = torch.zeros((h, w, c), dtype=dtype)
    img[i, j, :] = 1.
    return img


class StageB:
    """params: dict name -> mlp dict (see oracle.networks) plus 'light' [h,w,3].
    lxyz [L,3] / lareas [L] default to gen_light_xyz(light_h, 2*light_h)
    (shape.py:59-77); `light_idx` optionally maps each light direction to an
    env-map pixel (L != h*w configs)."""

    def __init__(self, params, config=None, light_h=16, lxyz=None, lareas=None,
                 light_idx=None, dtype=torch.float32):
        self.cfg = dict(DEFAULT_CONFIG)
        if config:
            self.cfg.update(config)
        self.p = params
        self.dtype = dtype
        if lxyz is None:
            lxyz, lareas = brdfmod.gen_light_xyz(light_h, 2 * light_h)
        # shape.py:75-76: cast to fp32 first (then to the compute dtype)
        self.lxyz = torch.as_tensor(
            np.asarray(lxyz, dtype=np.float32).reshape(-1, 3)).to(dtype)
        self.lareas = torch.as_tensor(
            np.asarray(lareas, dtype=np.float32).reshape(-1)).to(dtype)
        self.light_idx = None if light_idx is None else torch.as_tensor(
            np.asarray(light_idx, dtype=np.int64))

    # ---- helpers -----------------------------------------------------------
    def _t(self, x):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x
                               ).to(self.dtype)

    @property
    def light(self):
        """nerfactor.py:367-375: clip(light, 0, inf)."""
        return torch.clamp(self._t(self.p['light']), min=0.)

    def light_flat(self, light):
        lf = self._t(light).reshape(-1, 3)
        if self.light_idx is not None:
            lf = lf[self.light_idx]
        return lf

    @staticmethod
    def chunk_apply(func, x, dim, chunk_size):
        """shape.py:184-194 (the chunking does not change values)."""
        n = x.shape[0]
        y = torch.zeros((n, dim), dtype=x.dtype)
        for i in range(0, n, chunk_size):
            end_i = min(n, i + chunk_size)
            y[i:end_i] = func(x[i:end_i])
        return y

    def calc_ldir(self, pts):
        """shape.py:128-135."""
        surf2l = self.lxyz.reshape(1, -1, 3) - pts[:, None, :]
        return safe_l2_normalize(surf2l, 2)

    @staticmethod
    def calc_vdir(cam_loc, pts):
        """shape.py:137-144."""
        return safe_l2_normalize(cam_loc - pts, 1)

    def _point_mlp(self, name, pts):
        cfg = self.cfg
        pts_scaled = cfg['xyz_scale'] * pts

        def chunk_func(surf):
            e = nets.embed(surf, cfg['n_freqs_xyz'])
            return nets.mlp_forward(
                self.p[name + '_out'], nets.mlp_forward(self.p[name + '_mlp'], e))

        out_dim = self.p[name + '_out']['layers'][-1][0].shape[1]
        return self.chunk_apply(chunk_func, pts_scaled, out_dim, cfg['mlp_chunk'])

    def pred_normal_at(self, pts, eps=1e-6):
        """shape.py:196-211 (+eps on the raw output; caller normalises)."""
        return self._point_mlp('normal', pts) + eps

    def pred_albedo_at(self, pts):
        """nerfactor.py:377-396."""
        a = self._point_mlp('albedo', pts)
        return self.cfg['albedo_slope'] * a + self.cfg['albedo_bias']

    def pred_brdf_at(self, pts):
        """nerfactor.py:398-411 (z_dim latent) / nerfactor_microfacet.py:108-114
        (scalar roughness; the sigmoid lives in params['brdf_z_out']['act'])."""
        return self._point_mlp('brdf_z', pts)

    def pred_lvis_at(self, pts, surf2l):
        """shape.py:213-237."""
        cfg = self.cfg
        pts_scaled = cfg['xyz_scale'] * pts
        n, n_lights = surf2l.shape[0], surf2l.shape[1]
        surf2l_flat = surf2l.reshape(-1, 3)
        surf_flat = pts_scaled[:, None, :].expand(n, n_lights, 3).reshape(-1, 3)

        def chunk_func(ss):
            surf, s2l = ss[:, :3], ss[:, 3:]
            e = torch.cat((nets.embed(surf, cfg['n_freqs_xyz']),
                           nets.embed(s2l, cfg['n_freqs_ldir'])), -1)
            return nets.mlp_forward(
                self.p['lvis_out'], nets.mlp_forward(self.p['lvis_mlp'], e))

        ss = torch.cat((surf_flat, surf2l_flat), 1)
        lvis_flat = self.chunk_apply(chunk_func, ss, 1, cfg['mlp_chunk'])
        return lvis_flat.reshape(n, n_lights)

    def eval_brdf_learned(self, pts2l, pts2c, normal, albedo, brdf_prop):
        """nerfactor.py:413-461."""
        cfg = self.cfg
        z = brdf_prop
        n, n_l = pts2l.shape[0], pts2l.shape[1]
        world2local = brdfmod.gen_world2local(normal)
        vdir = torch.einsum('jkl,jl->jk', world2local, pts2c)
        ldir = torch.einsum('jkl,jnl->jnk', world2local, pts2l)
        ldir_flat = ldir.reshape(-1, 3)
        vdir_flat = vdir[:, None, :].expand(n, n_l, 3).reshape(-1, 3)
        rusink = brdfmod.dir2rusink(ldir_flat, vdir_flat)
        z_flat = z[:, None, :].expand(n, n_l, z.shape[1]).reshape(-1, z.shape[1])
        front_lit = ldir_flat[:, 2] > 0            # nerfactor.py:429-432
        rusink_fl, z_fl = rusink[front_lit], z_flat[front_lit]

        def chunk_func(rusink_z):
            r, zz = rusink_z[:, :3], rusink_z[:, 3:]
            e = nets.embed(r, cfg['n_freqs_rusink'])
            z_r = torch.cat((zz, e), dim=1)
            return nets.mlp_forward(
                self.p['brdf_out'], nets.mlp_forward(self.p['brdf_mlp'], z_r))

        rusink_z = torch.cat((rusink_fl, z_fl), 1)
        brdf_fl = self.chunk_apply(chunk_func, rusink_z, 1, cfg['mlp_chunk'])
        brdf_flat = torch.zeros((n * n_l, 1), dtype=pts2l.dtype)
        brdf_flat[front_lit] = brdf_fl
        spec = brdf_flat.reshape(n, n_l, 1).repeat(1, 1, 3)
        return albedo[:, None, :] / math.pi + spec * cfg['learned_brdf_scale']

    def eval_brdf_microfacet(self, pts2l, pts2c, normal, albedo, brdf_prop):
        """nerfactor_microfacet.py:116-124."""
        mf = brdfmod.Microfacet(f0=self.cfg['fresnel_f0'])
        return mf(pts2l, pts2c, normal, albedo=albedo, rough=brdf_prop)

    def eval_brdf_at(self, pts2l, pts2c, normal, albedo, brdf_prop):
        if self.cfg['brdf'] == 'microfacet':
            return self.eval_brdf_microfacet(pts2l, pts2c, normal, albedo, brdf_prop)
        return self.eval_brdf_learned(pts2l, pts2c, normal, albedo, brdf_prop)

    def render(self, light_vis, brdf, l, n, lights=None):
        """nerfactor.py:315-365: returns rgb [N,3] under self.light and, for each
        extra env-map in `lights` (list of [h,w,3]), rgb_relit [N,E,3]."""
        cos = torch.einsum('ijk,ik->ij', l, n)
        areas = self.lareas.reshape(1, -1, 1)
        front_lit = (cos > 0).to(cos.dtype)
        lvis = front_lit * light_vis

        def integrate(light):
            light_flat = self.light_flat(light)
            lt = lvis[:, :, None] * light_flat[None, :, :]
            contrib = brdf * lt * cos[:, :, None] * areas
            rgb = torch.sum(contrib, dim=1)
            rgb = torch.clamp(rgb, 0., 1.)
            if self.cfg['linear2srgb']:
                rgb = linear2srgb(rgb)
            return rgb

        rgb = integrate(self.light)
        rgb_relit = None
        if lights:
            rgb_relit = torch.cat(
                [integrate(x)[:, None, :] for x in lights], dim=1)
        return rgb, rgb_relit

    def novel_olat(self, light_res):
        """nerfactor.py:71-84: OLAT env-maps in row-major (i, j) order."""
        h, w = light_res
        ambient = (self.cfg['ambient_inten'] if self.cfg['white_bg'] else 0.) * \
            torch.ones((h, w, 3), dtype=self.dtype)
        return [self.cfg['olat_inten'] * one_hot_img(h, w, 3, i, j, self.dtype)
                + ambient for i in range(h) for j in range(w)]

    # ---- Model.call ---------------------------------------------------------
    def call(self, batch, mode='test', xyz_noise=None, relight_lights=None,
             albedo_scales=None, albedo_override=None, brdf_z_override=None):
        """nerfactor.py:181-313.  batch = (id_, hw, rayo, rayd, rgb, alpha, xyz,
        normal, lvis) as arrays; xyz_noise replaces tf.random.normal (:199) and
        must have the compacted (foreground) shape or be None."""
        if mode not in ('train', 'vali', 'test'):
            raise ValueError(mode)                     # models/base.py:107-110
        cfg = self.cfg
        _, _, rayo, _, rgb, alpha, xyz, normal, lvis = batch
        rayo, rgb, alpha, xyz, normal, lvis = [
            self._t(x) for x in (rayo, rgb, alpha, xyz, normal, lvis)]
        mask = alpha[:, 0] > 0
        rayo_m, rgb_m, xyz_m, normal_m, lvis_m = [
            x[mask] for x in (rayo, rgb, xyz, normal, lvis)]
        surf2l = self.calc_ldir(xyz_m)
        surf2c = self.calc_vdir(rayo_m, xyz_m)
        if xyz_noise is not None:
            xyz_noise = self._t(xyz_noise)
        # normals (nerfactor.py:203-214)
        if cfg['shape_mode'] == 'nerf':
            normal_pred, normal_jitter = normal_m, None
        else:
            normal_pred = self.pred_normal_at(xyz_m)
            normal_jitter = None if xyz_noise is None else \
                self.pred_normal_at(xyz_m + xyz_noise)
        normal_pred = safe_l2_normalize(normal_pred, 1)
        if normal_jitter is not None:
            normal_jitter = safe_l2_normalize(normal_jitter, 1)
        # light visibility (nerfactor.py:217-226)
        if cfg['shape_mode'] == 'nerf':
            lvis_pred, lvis_jitter = torch.clamp(lvis_m, 1e-8, 1.), None
        else:
            lvis_pred = self.pred_lvis_at(xyz_m, surf2l)
            lvis_jitter = None if xyz_noise is None else \
                self.pred_lvis_at(xyz_m + xyz_noise, surf2l)
        # albedo (nerfactor.py:228-242)
        albedo = self.pred_albedo_at(xyz_m)
        albedo_jitter = None if xyz_noise is None else \
            self.pred_albedo_at(xyz_m + xyz_noise)
        if albedo_scales is not None:
            albedo = self._t(albedo_scales).reshape(1, 3) * albedo
        if albedo_override is not None:
            ao = self._t(albedo_override)
            albedo = ao[None, :].expand(albedo.shape[0], 3) if ao.dim() == 1 \
                else ao[mask]
        # BRDF property (nerfactor.py:244-260)
        brdf_prop = self.pred_brdf_at(xyz_m)
        brdf_prop_jitter = None if xyz_noise is None else \
            self.pred_brdf_at(xyz_m + xyz_noise)
        if cfg['normalize_brdf_z']:
            brdf_prop = safe_l2_normalize(brdf_prop, 1)
            if brdf_prop_jitter is not None:
                brdf_prop_jitter = safe_l2_normalize(brdf_prop_jitter, 1)
        if brdf_z_override is not None:
            zo = self._t(brdf_z_override).reshape(1, -1)
            brdf_prop = zo.expand(brdf_prop.shape[0], zo.shape[1])
        brdf = self.eval_brdf_at(surf2l, surf2c, normal_pred, albedo, brdf_prop)
        rgb_pred, rgb_relit = self.render(
            lvis_pred, brdf, surf2l, normal_pred, lights=relight_lights)

        # scatter back to the full ray set (nerfactor.py:268-293)
        n = alpha.shape[0]

        def scatter(v):
            if v is None:
                return None
            out = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
            out[mask] = v
            return out

        pred = {'rgb': scatter(rgb_pred), 'normal': scatter(normal_pred),
                'lvis': scatter(lvis_pred), 'albedo': scatter(albedo),
                'brdf': scatter(brdf_prop)}
        if rgb_relit is not None:
            pred['rgb_relit'] = scatter(rgb_relit)
        gt = {'rgb': scatter(rgb_m), 'normal': scatter(normal_m),
              'lvis': scatter(lvis_m), 'alpha': alpha}
        loss_kwargs = {
            'mode': mode, 'normal_jitter': scatter(normal_jitter),
            'lvis_jitter': scatter(lvis_jitter),
            'brdf_prop_jitter': scatter(brdf_prop_jitter),
            'albedo_jitter': scatter(albedo_jitter)}
        return pred, gt, loss_kwargs

    # ---- compute_loss -------------------------------------------------------
    def compute_loss(self, pred, gt, mode, normal_jitter, lvis_jitter,
                     brdf_prop_jitter, albedo_jitter, weights=None):
        """nerfactor.py:463-541 -> per-ray loss [N]."""
        w = {'normal_loss_weight': 0.1, 'lvis_loss_weight': 0.1,
             'normal_smooth_weight': 0.05, 'lvis_smooth_weight': 0.05,
             'albedo_smooth_weight': 0.05, 'brdf_smooth_weight': 0.01,
             'smooth_use_l1': True, 'light_tv_weight': 5e-6,
             'light_achro_weight': 0.0}       # nerfactor.ini:54-75
        if weights:
            w.update(weights)
        mse = lambda a, b: torch.mean((a - b) ** 2, dim=-1)     # keras MSE
        mae = lambda a, b: torch.mean(torch.abs(a - b), dim=-1)  # keras MAE
        smooth = mae if w['smooth_use_l1'] else mse
        alpha = gt['alpha']
        bgv = 1. if self.cfg['white_bg'] else 0.

        def blend(x):
            return alpha_blend(x, alpha, torch.full_like(x, bgv))

        rgb_pred, rgb_gt = blend(pred['rgb']), blend(gt['rgb'])
        normal_pred, normal_gt = blend(pred['normal']), blend(gt['normal'])
        lvis_pred, lvis_gt = blend(pred['lvis']), blend(gt['lvis'])
        loss = mse(rgb_gt, rgb_pred)
        if mode == 'vali':
            return loss
        if self.cfg['shape_mode'] in ('scratch', 'finetune'):
            loss = loss + w['normal_loss_weight'] * mse(normal_gt, normal_pred)
            loss = loss + w['lvis_loss_weight'] * mse(lvis_gt, lvis_pred)
            if normal_jitter is not None:
                loss = loss + w['normal_smooth_weight'] * smooth(
                    normal_pred, normal_jitter)
            if lvis_jitter is not None:
                loss = loss + w['lvis_smooth_weight'] * smooth(
                    lvis_pred, lvis_jitter)
        if albedo_jitter is not None:
            loss = loss + w['albedo_smooth_weight'] * smooth(
                pred['albedo'], albedo_jitter)
        if brdf_prop_jitter is not None:
            loss = loss + w['brdf_smooth_weight'] * smooth(
                pred['brdf'], brdf_prop_jitter)
        if mode == 'train':
            light = self.light
            if w['light_tv_weight'] > 0:
                dx = light - torch.roll(light, 1, 1)
                dy = light - torch.roll(light, 1, 0)
                loss = loss + w['light_tv_weight'] * torch.sum(dx ** 2 + dy ** 2)
            if w['light_achro_weight'] > 0:
                dc = light - torch.roll(light, 1, 2)
                loss = loss + w['light_achro_weight'] * torch.sum(dc ** 2)
        return loss

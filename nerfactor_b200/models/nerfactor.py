"""Mirror of nerfactor/models/nerfactor.py: the joint render-and-relight model.

`Model.call(batch, mode, relight_olat, relight_probes, albedo_scales,
albedo_override, brdf_z_override) -> (pred, gt, loss_kwargs, to_vis)` with the
reference's 9-tuple batch (nerfactor/datasets/nerf_shape.py:72-95).  Everything
per-(point, light) runs in the fused kernels of libnerfactor_b200.so; torch ops
here only compact / scatter the foreground rays and do [N,3]-sized glue.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib
from ..config import default_config
from ..networks import mlp
from ..networks.embedder import Embedder
from ..util import math as mathutil, img as imgutil, io as ioutil, config as configutil, \
    light as lightutil
from ..util.io import restore_model
from .shape import Model as ShapeModel, to_device
from .brdf import Model as BRDFModel
from ._visualize import NeRFactorVis


class Model(NeRFactorVis, ShapeModel):
    def __init__(self, config, debug=False, params=None, ctx=None, precision='f16',
                 config_brdf=None, test_time_jitter=False, allow_uninitialised_prior=False):
        # BRDF (nerfactor.py:36-42): the prior's config sits next to its checkpoint
        # (`brdf_model_ckpt` -> `<outroot>/<xname>.ini`).  Like the reference, a configured
        # checkpoint (or its .ini) that does not exist is an error -- a typo must not train or
        # render with a random frozen prior.  Weights handed over explicitly (`params`, tests and
        # the benchmark) or `allow_uninitialised_prior=True` are the only ways around a checkpoint.
        self.brdf_model_ckpt = config.get('DEFAULT', 'brdf_model_ckpt', fallback='')
        self._explicit_weights = params is not None or allow_uninitialised_prior
        if config_brdf is None and self.brdf_model_ckpt and self._uses_brdf_prior:
            ini = configutil.get_config_ini(self.brdf_model_ckpt)
            if os.path.exists(ini):
                config_brdf = ioutil.read_config(ini)
            elif not self._explicit_weights:
                raise FileNotFoundError(
                    "config of the BRDF prior not found: %s (from brdf_model_ckpt = %s)"
                    % (ini, self.brdf_model_ckpt))
        self.config_brdf = config_brdf or default_config('brdf')
        self.pred_brdf = config.getboolean('DEFAULT', 'pred_brdf')
        if not self.pred_brdf:
            # nerfactor.py:256 calls an undefined _get_default_brdf_at: dead branch
            raise NotImplementedError("pred_brdf=False is dead code in the reference")
        self._init_brdf_dims()
        self.shape_mode = config.get('DEFAULT', 'shape_mode')
        if self.shape_mode not in ('scratch', 'frozen', 'finetune', 'nerf'):
            raise ValueError(self.shape_mode)
        # Reference quirk kept as an option: jittered copies of every network are
        # evaluated even in test mode (nerfactor.py:198-232) although they only
        # feed the loss.  Default: skip them at test time.
        self.test_time_jitter = test_time_jitter
        super().__init__(config, debug=debug, params=None, ctx=ctx, precision=precision)
        self.albedo_smooth_weight = config.getfloat('DEFAULT', 'albedo_smooth_weight')
        self.brdf_smooth_weight = config.getfloat('DEFAULT', 'brdf_smooth_weight')
        self._init_brdf_model(params)
        # Lighting (nerfactor.py:62-66, 367-375)
        light_h = self.config.getint('DEFAULT', 'light_h')
        self.light_res = (light_h, 2 * light_h)
        maxv = self.config.getfloat('DEFAULT', 'light_init_max')
        self._light = torch.rand(self.light_res + (3,), device=self.device) * maxv
        self.light_idx = None            # optional light-direction -> env-map pixel map
        # Novel lighting for relighting at test time (nerfactor.py:67-92)
        self.olat_inten = self.config.getfloat('DEFAULT', 'olat_inten', fallback=200)
        ambi = self.config.getfloat('DEFAULT', 'ambient_inten', fallback=0)
        self.ambient_inten = ambi if self.white_bg else 0.
        self.novel_olat = _LazyOlat(self)
        # (2) light probes: every .hdr of `test_envmap_dir` resampled to the light resolution
        # (nerfactor.py:85-92, 169-179); name -> [h, 2h, 3].  Callers may also fill it directly.
        self.novel_probes = OrderedDict()
        envmap_dir = self.config.get('DEFAULT', 'test_envmap_dir', fallback='')
        for name, envmap in lightutil.load_probes(envmap_dir, light_h).items():
            self.novel_probes[name] = to_device(envmap, self.device)
        self._restore_submodels()
        if params is not None:
            self.load_params(params)

    # ------------------------------------------------------------ construction
    _uses_brdf_prior = True          # the microfacet variant has no learned prior

    def _init_brdf_dims(self):
        self.z_dim = self.config_brdf.getint('DEFAULT', 'z_dim')
        self.normalize_brdf_z = self.config_brdf.getboolean('DEFAULT', 'normalize_z')

    def _init_brdf_model(self, params):
        self.brdf_model = BRDFModel(self.config_brdf, params=params if params and
                                    'brdf_mlp' in params else None)

    def _load_light(self, path):
        """nerfactor.py:169-179."""
        return to_device(lightutil.load_light(path, self.light_res[0]), self.device)

    def _restore_submodels(self):
        """What the reference restores while constructing the model, when the checkpoints named in
        the config exist: the frozen BRDF prior (nerfactor.py:57-60) and, for shape_mode
        'frozen' / 'finetune', the pre-trained normal / visibility MLPs (nerfactor.py:156-163)."""
        from ..util import tfckpt

        def have(ckpt, what):
            if ckpt and os.path.exists(ckpt + '.index'):
                return True
            if self._explicit_weights:
                return False               # weights come from `params` / explicit opt-out
            raise FileNotFoundError(
                "%s checkpoint not found: %r (.index missing); pass `params=` or "
                "`allow_uninitialised_prior=True` to construct the model without it" % (what, ckpt))
        if self.brdf_model is not None and have(self.brdf_model_ckpt, 'BRDF-prior (brdf_model_ckpt)'):
            restore_model(self.brdf_model, self.brdf_model_ckpt)
            n_z = np.asarray(self.brdf_model.latent_code.z).shape[0]
            if len(self.brdf_model.brdf_names) != n_z:
                self.brdf_model.brdf_names = ['brdf_%03d' % i for i in range(n_z)]
        ckpt = self.config.get('DEFAULT', 'shape_model_ckpt', fallback='')
        if self.shape_mode in ('frozen', 'finetune') and have(ckpt, 'shape (shape_model_ckpt)'):
            params = tfckpt.params_from_checkpoint(ckpt)
            ShapeModel.load_params(self, {k: params[k] for k in (
                'normal_mlp', 'normal_out', 'lvis_mlp', 'lvis_out')})

    def _init_embedder(self):
        """nerfactor.py:107-126."""
        embedder = super()._init_embedder()
        n = self.config_brdf.getint('DEFAULT', 'n_freqs')
        embedder['rusink'] = Embedder(incl_input=True, in_dims=3, log2_max_freq=n - 1,
                                      n_freqs=n, log_sampling=True)
        return embedder

    def _init_net(self):
        """nerfactor.py:128-167 (the restored shape networks are plain members)."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'mlp_depth')
        s = self.config.getint('DEFAULT', 'mlp_skip_at')
        trunk = lambda: mlp.Network([w] * d, act=['relu'] * d, skip_at=[s])
        net = {}
        net['albedo_mlp'] = trunk()
        net['albedo_out'] = mlp.Network([3], act=['sigmoid'])
        net['brdf_z_mlp'] = trunk()
        net['brdf_z_out'] = mlp.Network([self.z_dim], act=None)
        if self.shape_mode != 'nerf':
            net['normal_mlp'] = trunk()
            net['normal_out'] = mlp.Network([3], act=None)
            net['lvis_mlp'] = trunk()
            net['lvis_out'] = mlp.Network([1], act=['sigmoid'])
            if self.shape_mode == 'frozen':
                for k in ('normal_mlp', 'normal_out', 'lvis_mlp', 'lvis_out'):
                    for layer in net[k].layers:
                        layer.trainable = False
        return net

    def load_params(self, params):
        super().load_params(params)
        if 'light' in params:
            self._light = to_device(params['light'], self.device)
        if 'brdf_mlp' in params and hasattr(self, 'brdf_model'):
            for k in ('brdf_mlp', 'brdf_out'):
                self.brdf_model.net[k].load(params[k])
            self._packed.pop('brdf', None)

    @property
    def light(self):
        """nerfactor.py:367-375: no negative light."""
        return torch.clamp(self._light, min=0.)

    def set_lights(self, lxyz, lareas, light_idx=None):
        super().set_lights(lxyz, lareas)
        self.light_idx = None if light_idx is None else to_device(
            np.asarray(light_idx, np.int32), self.device, torch.int32)

    # ------------------------------------------------------------ network evals
    def _pred_albedo_at(self, pts):
        """nerfactor.py:377-396."""
        albedo_scale = self.config.getfloat('DEFAULT', 'albedo_slope', fallback=0.7)
        albedo_bias = self.config.getfloat('DEFAULT', 'albedo_bias', fallback=0.1)
        return albedo_scale * self._pred_point('albedo', pts) + albedo_bias

    def _pred_brdf_at(self, pts):
        """nerfactor.py:398-411."""
        return self._pred_point('brdf_z', pts)

    def _packed_brdf_mlp(self):
        if 'brdf' not in self._packed:
            trunk, head = self.brdf_model.net['brdf_mlp'], self.brdf_model.net['brdf_out']
            self._packed['brdf'] = _lib.PackedMlp(
                self.ctx, 'brdf', trunk.weights() + head.weights(), trunk.skip_at[0],
                'softplus', n_freqs_a=self.embedder['rusink'].n_freqs, z_dim=self.z_dim)
        return self._packed['brdf']

    def _eval_brdf_at(self, pts2l, pts2c, normal, albedo, brdf_prop, pts=None, cam=None):
        """nerfactor.py:413-461.  Returns the achromatic specular lobe spec[N,L]
        (the Lambertian term albedo/pi and learned_brdf_scale are applied inside the
        rendering-equation kernel, nerfactor.py:460).  `pts2l` / `pts2c` are accepted
        for signature parity; directions are rebuilt on chip from pts / cam."""
        spec = _lib.brdf_learned_fwd(
            self.ctx, self._packed_brdf_mlp(), pts, normal, cam, brdf_prop.contiguous(),
            self.lxyz.reshape(-1, 3), self.precision)
        return {'spec': spec}

    def _render(self, light_vis, brdf, l, n, relight_olat=False, relight_probes=False,
                pts=None, cam=None, albedo=None):
        """nerfactor.py:315-365: rgb under the learned light, then under every OLAT
        and every probe.  `brdf` is what `_eval_brdf_at` returned."""
        linear2srgb = self.config.getboolean('DEFAULT', 'linear2srgb')
        common = dict(lxyz=self.lxyz.reshape(-1, 3), lareas=self.lareas.reshape(-1),
                      linear2srgb=linear2srgb, **self._brdf_kernel_args(brdf))
        lights = [self.light.reshape(-1, 3)]
        if relight_probes:
            lights += [to_device(v, self.device).reshape(-1, 3)
                       for v in self.novel_probes.values()]
        light = torch.stack(lights, 0).contiguous()
        rgb_all = _lib.integrate_fwd(self.ctx, pts, n, cam, albedo, light_vis,
                                     light=light, light_idx=self.light_idx, **common)
        rgb = rgb_all[:, 0, :]
        rgb_probes = rgb_all[:, 1:, :] if relight_probes else None
        rgb_olat = None
        if relight_olat:
            rgb_olat = _lib.integrate_olat_fwd(
                self.ctx, pts, n, cam, albedo, light_vis, olat_inten=self.olat_inten,
                ambient=self.ambient_inten, **common)
        return rgb, rgb_olat, rgb_probes

    def _brdf_kernel_args(self, brdf):
        return {'spec': brdf['spec'],
                'spec_scale': self.config.getfloat('DEFAULT', 'learned_brdf_scale')}

    # ------------------------------------------------------------------- call
    def call(self, batch, mode='train', relight_olat=False, relight_probes=False,
             albedo_scales=None, albedo_override=None, brdf_z_override=None,
             xyz_noise=None):
        """nerfactor.py:181-313.  `xyz_noise` (compacted [N_fg,3]) replaces the
        reference's tf.random.normal (:199) when given."""
        xyz_jitter_std = self.config.getfloat('DEFAULT', 'xyz_jitter_std')
        self._validate_mode(mode)
        id_, hw, rayo, _, rgb, alpha, xyz, normal, lvis = batch
        dev = self.device
        alpha = to_device(alpha, dev)
        rayo, rgb, xyz, normal = [to_device(x, dev) for x in (rayo, rgb, xyz, normal)]
        need_lvis_gt = lvis is not None
        if need_lvis_gt:
            lvis = to_device(lvis, dev)
        # Mask out 100% background (:188-193)
        mask = alpha[:, 0] > 0
        ind = torch.nonzero(mask, as_tuple=False)[:, 0]
        sel = lambda x: x.index_select(0, ind).contiguous()
        rayo_m, rgb_m, xyz_m, normal_m = sel(rayo), sel(rgb), sel(xyz), sel(normal)
        lvis_m = sel(lvis) if need_lvis_gt else None
        # Jitter (:198-201)
        jitter = xyz_jitter_std > 0 and (mode != 'test' or self.test_time_jitter)
        if xyz_noise is not None:
            xyz_noise = to_device(xyz_noise, dev)
        elif jitter:
            xyz_noise = torch.randn_like(xyz_m) * xyz_jitter_std
        xyz_j = None if xyz_noise is None else (xyz_m + xyz_noise).contiguous()
        # Normals (:203-214)
        if self.shape_mode == 'nerf':
            normal_pred, normal_jitter = normal_m, None
        else:
            normal_pred = self._pred_normal_at(xyz_m)
            normal_jitter = None if xyz_j is None else self._pred_normal_at(xyz_j)
        normal_pred = mathutil.safe_l2_normalize(normal_pred, axis=1)
        if normal_jitter is not None:
            normal_jitter = mathutil.safe_l2_normalize(normal_jitter, axis=1)
        # Light visibility (:217-226)
        if self.shape_mode == 'nerf':
            lvis_pred, lvis_jitter = torch.clamp(lvis_m, 1e-8, 1.), None
        else:
            lvis_pred = self._pred_lvis_at(xyz_m)
            lvis_jitter = None if xyz_j is None else self._pred_lvis_jitter_at(xyz_j, xyz_m)
        # Albedo (:228-242)
        albedo = self._pred_albedo_at(xyz_m)
        albedo_jitter = None if xyz_j is None else self._pred_albedo_at(xyz_j)
        if albedo_scales is not None:
            albedo = to_device(albedo_scales, dev).reshape(1, 3) * albedo
        if albedo_override is not None:
            ao = to_device(albedo_override, dev)
            albedo = ao[None, :].expand(albedo.shape[0], 3) if ao.dim() == 1 else sel(ao)
        albedo = albedo.contiguous()
        # BRDF property (:244-260)
        brdf_prop = self._pred_brdf_at(xyz_m)
        brdf_prop_jitter = None if xyz_j is None else self._pred_brdf_at(xyz_j)
        if self.normalize_brdf_z:
            brdf_prop = mathutil.safe_l2_normalize(brdf_prop, axis=1)
            if brdf_prop_jitter is not None:
                brdf_prop_jitter = mathutil.safe_l2_normalize(brdf_prop_jitter, axis=1)
        if brdf_z_override is not None:
            zo = to_device(brdf_z_override, dev).reshape(1, self.z_dim)
            brdf_prop = zo.expand(brdf_prop.shape[0], self.z_dim).contiguous()
        normal_pred = normal_pred.contiguous()
        brdf = self._eval_brdf_at(None, None, normal_pred, albedo, brdf_prop,
                                  pts=xyz_m, cam=rayo_m)
        # Rendering equation (:264-266)
        rgb_pred, rgb_olat, rgb_probes = self._render(
            lvis_pred.contiguous(), brdf, None, normal_pred, relight_olat=relight_olat,
            relight_probes=relight_probes, pts=xyz_m, cam=rayo_m, albedo=albedo)
        # Put values back into the full shape (:268-293)
        n = alpha.shape[0]

        def scatter(v):
            if v is None:
                return None
            out = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            out.index_copy_(0, ind, v)
            return out

        pred = {'rgb': scatter(rgb_pred), 'normal': scatter(normal_pred),
                'lvis': scatter(lvis_pred), 'albedo': scatter(albedo),
                'brdf': scatter(brdf_prop)}
        if rgb_olat is not None:
            pred['rgb_olat'] = scatter(rgb_olat)
        if rgb_probes is not None:
            pred['rgb_probes'] = scatter(rgb_probes)
        gt = {'rgb': scatter(rgb_m), 'normal': scatter(normal_m),
              'lvis': scatter(lvis_m), 'alpha': alpha}
        loss_kwargs = {
            'mode': mode, 'normal_jitter': scatter(normal_jitter),
            'lvis_jitter': scatter(lvis_jitter),
            'brdf_prop_jitter': scatter(brdf_prop_jitter),
            'albedo_jitter': scatter(albedo_jitter)}
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    # ------------------------------------------------------------ fused rendering
    def render_rgb(self, batch, relight_probes=False, want_lvis=False, all_lights=False):
        """The test-mode RGB of `call` (nerfactor.py:181-313: no jitter, no edits) through the
        fused Stage-B op nf_stageB_fused_fwd: per-point networks, then light-visibility network ->
        BRDF -> rendering equation in one call, the [N, L] light-visibility tensor not
        materialised (unless `want_lvis`).  When it is not an output, the visibility network only
        runs on the lights facing the shading normal -- the renderer multiplies the visibility of
        all others by zero (nerfactor.py:329-330), so 'rgb' does not change; `all_lights=True`
        evaluates it for every light anyway.  Returns a dict with 'rgb', 'normal', 'albedo', 'brdf'
        (+ 'rgb_probes', 'lvis') in the full ray shape, background rows zero -- the same values
        `call(batch, 'test')` returns for those keys."""
        if self.shape_mode == 'nerf':
            raise NotImplementedError("shape_mode = 'nerf' reads lvis from the batch: use call()")
        _, _, rayo, _, _, alpha, xyz, _, _ = batch
        dev = self.device
        alpha, rayo, xyz = [to_device(x, dev) for x in (alpha, rayo, xyz)]
        ind = torch.nonzero(alpha[:, 0] > 0, as_tuple=False)[:, 0]
        sel = lambda x: x.index_select(0, ind).contiguous()
        rayo_m, xyz_m = sel(rayo), sel(xyz)
        normal = mathutil.safe_l2_normalize(self._pred_normal_at(xyz_m), axis=1).contiguous()
        albedo = self._pred_albedo_at(xyz_m).contiguous()
        brdf_prop = self._pred_brdf_at(xyz_m)
        if self.normalize_brdf_z:
            brdf_prop = mathutil.safe_l2_normalize(brdf_prop, axis=1)
        brdf_prop = brdf_prop.contiguous()
        lights = [self.light.reshape(-1, 3)]
        if relight_probes:
            lights += [to_device(v, dev).reshape(-1, 3) for v in self.novel_probes.values()]
        light = torch.stack(lights, 0).contiguous()
        rgb_all, lvis = self._fused_stage_b(xyz_m, normal, rayo_m, albedo, brdf_prop, light,
                                            want_lvis, all_lights)
        n = alpha.shape[0]

        def scatter(v):
            out = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            out.index_copy_(0, ind, v)
            return out
        out = {'rgb': scatter(rgb_all[:, 0, :]), 'normal': scatter(normal),
               'albedo': scatter(albedo), 'brdf': scatter(brdf_prop)}
        if relight_probes:
            out['rgb_probes'] = scatter(rgb_all[:, 1:, :])
        if want_lvis:
            out['lvis'] = scatter(lvis)
        return out

    def _fused_stage_b(self, pts, normal, cam, albedo, brdf_prop, light, want_lvis, all_lights=False):
        """Learned-MERL lobe: latent z + the BRDF prior's network."""
        m_lvis = self._packed_mlp('lvis', 'lvis', n_freqs_a=self.embedder['xyz'].n_freqs,
                                  n_freqs_b=self.embedder['ldir'].n_freqs)
        return _lib.stageB_fused_fwd(
            self.ctx, m_lvis, pts, normal, cam, albedo, self.lxyz, self.lareas, light,
            z=brdf_prop, mlp_brdf=self._packed_brdf_mlp(), light_idx=self.light_idx,
            spec_scale=self.config.getfloat('DEFAULT', 'learned_brdf_scale'),
            xyz_scale=self.xyz_scale,
            linear2srgb=self.config.getboolean('DEFAULT', 'linear2srgb'),
            precision=self.precision, want_lvis=want_lvis, all_lights=all_lights)

    # ------------------------------------------------------------------- loss
    def compute_loss(self, pred, gt, **kwargs):
        """nerfactor.py:463-541 -> per-ray loss [N] (keys popped like the reference)."""
        cfg = self.config
        normal_loss_weight = cfg.getfloat('DEFAULT', 'normal_loss_weight')
        lvis_loss_weight = cfg.getfloat('DEFAULT', 'lvis_loss_weight')
        smooth_use_l1 = cfg.getboolean('DEFAULT', 'smooth_use_l1')
        light_tv_weight = cfg.getfloat('DEFAULT', 'light_tv_weight')
        light_achro_weight = cfg.getfloat('DEFAULT', 'light_achro_weight')
        mse = lambda a, b: torch.mean((a - b) ** 2, dim=-1)
        mae = lambda a, b: torch.mean(torch.abs(a - b), dim=-1)
        smooth_loss = mae if smooth_use_l1 else mse
        mode = kwargs.pop('mode')
        normal_jitter = kwargs.pop('normal_jitter')
        lvis_jitter = kwargs.pop('lvis_jitter')
        albedo_jitter = kwargs.pop('albedo_jitter')
        brdf_prop_jitter = kwargs.pop('brdf_prop_jitter')
        alpha = gt['alpha']
        bgv = 1. if self.white_bg else 0.
        blend = lambda x: imgutil.alpha_blend(x, alpha, torch.full_like(x, bgv))
        rgb_pred, rgb_gt = blend(pred['rgb']), blend(gt['rgb'])
        normal_pred, normal_gt = blend(pred['normal']), blend(gt['normal'])
        lvis_pred, lvis_gt = blend(pred['lvis']), blend(gt['lvis'])
        loss = mse(rgb_gt, rgb_pred)
        if mode == 'vali':
            return loss
        if self.shape_mode in ('scratch', 'finetune'):
            loss = loss + normal_loss_weight * mse(normal_gt, normal_pred)
            loss = loss + lvis_loss_weight * mse(lvis_gt, lvis_pred)
            if normal_jitter is not None:
                loss = loss + self.normal_smooth_weight * smooth_loss(normal_pred, normal_jitter)
            if lvis_jitter is not None:
                loss = loss + self.lvis_smooth_weight * smooth_loss(lvis_pred, lvis_jitter)
        if albedo_jitter is not None:
            loss = loss + self.albedo_smooth_weight * smooth_loss(pred['albedo'], albedo_jitter)
        if brdf_prop_jitter is not None:
            loss = loss + self.brdf_smooth_weight * smooth_loss(pred['brdf'], brdf_prop_jitter)
        if mode == 'train':
            light = self.light
            if light_tv_weight > 0:
                dx = light - torch.roll(light, 1, 1)
                dy = light - torch.roll(light, 1, 0)
                loss = loss + light_tv_weight * torch.sum(dx ** 2 + dy ** 2)
            if light_achro_weight > 0:
                dc = light - torch.roll(light, 1, 2)
                loss = loss + light_achro_weight * torch.sum(dc ** 2)
        return loss


class _LazyOlat:
    """The 512 OLAT env-maps of nerfactor.py:71-84, generated on demand (the kernel
    never needs them materialised: nf_integrate_olat_fwd takes intensity + ambient)."""

    def __init__(self, model):
        self.m = model

    def __len__(self):
        return self.m.light_res[0] * self.m.light_res[1]

    def keys(self):
        h, w = self.m.light_res
        return ['%04d-%04d' % (i, j) for i in range(h) for j in range(w)]

    def __getitem__(self, key):
        i, j = [int(x) for x in key.split('-')]
        h, w = self.m.light_res
        env = torch.full((h, w, 3), self.m.ambient_inten, device=self.m.device)
        env[i, j, :] += self.m.olat_inten
        return env

    def items(self):
        return ((k, self[k]) for k in self.keys())

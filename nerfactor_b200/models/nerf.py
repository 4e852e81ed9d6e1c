"""Mirror of nerfactor/models/nerf.py (forward): the coarse / fine networks
(`net['coarse_enc']`, `net['coarse_sigma_out']`, `net['coarse_bottleneck']`,
`net['coarse_rgb_out']`, ... nerf.py:53-71), the `xyz` / `view` embedders, the static
samplers `gen_z`, `gen_z_fine`, `accumulate_sigma` (nerf.py:120-147, 184-212) and the colour
rendering `_render_rays` / `_accumulate` / `_eval_nerf_at` / `call` (nerf.py:100-118, 149-290,
SURVEY.md 8f.2).  NeRF's own training: trainvali.NerfTrainer."""
import numpy as np
import torch

from .. import _lib
from ..networks import mlp
from ..networks.embedder import Embedder
from .base import Model as BaseModel
from ._visualize import NerfVis


class Model(NerfVis, BaseModel):
    def __init__(self, config, debug=False, params=None, ctx=None, precision='f16'):
        super().__init__(config, debug=debug)
        self.ctx = ctx or _lib.default_context()
        self.device = self.ctx.device
        self.precision = precision
        self.use_views = self.config.getboolean('DEFAULT', 'use_views')
        self.near = self.config.getfloat('DEFAULT', 'near')
        self.far = self.config.getfloat('DEFAULT', 'far')
        self.n_samples_fine = self.config.getint('DEFAULT', 'n_samples_fine')
        self.white_bg = self.config.getboolean('DEFAULT', 'white_bg')
        self.embedder = self._init_embedder()
        self.net = {}
        for k, v in self._init_net().items():
            self.net['coarse_' + k] = v
        if self.n_samples_fine > 0:
            for k, v in self._init_net().items():
                self.net['fine_' + k] = v
        rng = np.random.default_rng(0)
        width = self.config.getint('DEFAULT', 'mlp_width')
        for pref in ('coarse_', 'fine_'):
            if pref + 'enc' in self.net:
                self.net[pref + 'enc'].build(self.embedder['xyz'].out_dims, rng)
                self.net[pref + 'sigma_out'].build(width, rng)
                if self.use_views:
                    self.net[pref + 'bottleneck'].build(width, rng)
                    self.net[pref + 'rgb_out'].build(width + self.embedder['view'].out_dims, rng)
        self._packed = {}
        if params is not None:
            self.load_params(params)

    def _init_net(self):
        """nerf.py:53-71."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'enc_depth')
        act = self.config.get('DEFAULT', 'act', fallback='relu')
        if act != 'relu':
            raise NotImplementedError(act)
        net = {'enc': mlp.Network([w] * d, act=[act] * d, skip_at=[d // 2])}
        if not self.use_views:
            raise NotImplementedError("use_views = False (rgbs_out head, nerf.py:61-65)")
        net['sigma_out'] = mlp.Network([1], act=[None])                 # ReLU later
        net['bottleneck'] = mlp.Network([w], act=[None])
        net['rgb_out'] = mlp.Network([w // 2, 3], act=[act, None])      # sigmoid later
        return net

    def _init_embedder(self):
        """nerf.py:73-98."""
        n = self.config.getint('DEFAULT', 'n_freqs_xyz')
        nv = self.config.getint('DEFAULT', 'n_freqs_view')
        mk = lambda k: Embedder(incl_input=True, in_dims=3, log2_max_freq=k - 1, n_freqs=k,
                                log_sampling=True)
        return {'xyz': mk(n), 'view': mk(nv)}

    def load_params(self, params):
        for k, net in self.net.items():
            if k in params:
                net.load(params[k])
        self._packed.clear()

    def packed_sigma(self, use_fine):
        pref = 'fine_' if use_fine else 'coarse_'
        if pref not in self._packed:
            trunk, head = self.net[pref + 'enc'], self.net[pref + 'sigma_out']
            self._packed[pref] = _lib.PackedMlp(
                self.ctx, 'sigma', trunk.weights() + head.weights(), trunk.skip_at[0], None,
                n_freqs_a=self.embedder['xyz'].n_freqs)
        return self._packed[pref]

    def packed_nerf(self, use_fine):
        """Sigma network + colour branch packed together (nf_mlp_attach_rgb)."""
        pref = 'fine_' if use_fine else 'coarse_'
        key = pref + 'rgb'
        if key not in self._packed:
            trunk, head = self.net[pref + 'enc'], self.net[pref + 'sigma_out']
            rgb = {'bottleneck': self.net[pref + 'bottleneck'].weights()[0],
                   'rgb_out': self.net[pref + 'rgb_out'].weights(),
                   'n_freqs_view': self.embedder['view'].n_freqs}
            self._packed[key] = _lib.PackedMlp(
                self.ctx, 'sigma', trunk.weights() + head.weights(), trunk.skip_at[0], None,
                n_freqs_a=self.embedder['xyz'].n_freqs, rgb=rgb)
        return self._packed[key]

    # ---- colour rendering (nerf.py:100-118, 149-290) -------------------------------
    def _eval_nerf_at(self, rayo, rayd, z, use_fine=False, precision=None):
        """nerf.py:254-290 at the samples pts = rayo + rayd z, views = rayd (never
        materialised; nerf.py:162-165): rgbs [n, S, 4] = (raw r, g, b, raw sigma).
        precision 'fp32': layer-by-layer FP32 Dense kernels on materialised activations (the
        tight-parity path); 'f16' / 'bf16': one fused tcgen05 kernel."""
        prec = precision or self.precision
        prec = {'f16e': 'f16'}.get(prec, prec)        # colour kernel: no split-encoding variant
        if prec != 'fp32':
            return _lib.nerf_fwd(self.ctx, self.packed_nerf(use_fine), rayo, rayd, z, prec)
        from .. import autodiff as ad
        pref = 'fine_' if use_fine else 'coarse_'
        n, S = z.shape
        dev = self.device
        lay = lambda name: [(torch.as_tensor(w).to(dev), torch.as_tensor(b).to(dev))
                            for w, b in self.net[pref + name].weights()]
        pts = (rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]).reshape(-1, 3)
        views = rayd[:, None, :].expand(n, S, 3).reshape(-1, 3)
        with torch.no_grad():
            enc = self.net[pref + 'enc']
            feat = ad.mlp_apply(ad.embed(pts, self.embedder['xyz'].n_freqs), lay('enc'),
                                ['relu'] * len(enc.layers), enc.skip_at)
            sigma = ad.mlp_apply(feat, lay('sigma_out'), [None], None)
            bott = ad.mlp_apply(feat, lay('bottleneck'), [None], None)
            fv = torch.cat((bott, ad.embed(views, self.embedder['view'].n_freqs)), -1)
            rgb = ad.mlp_apply(fv, lay('rgb_out'), ['relu', None], None)
        return torch.cat((rgb, sigma), -1).reshape(n, S, 4)

    def _accumulate(self, rgbs, z, rayd, eps=1e-10):
        """nerf.py:214-252 -> (rgb, occu, depth, disp, weights)."""
        sigma = rgbs[:, :, 3].contiguous()
        rgb_s = torch.sigmoid(rgbs[:, :, :3]).contiguous()
        weights, occu, depth, _, rgb = _lib.composite(self.ctx, sigma, z, rayd, rayd,
                                                      normal=rgb_s, want_surf=False)
        disp = 1. / torch.clamp(depth, min=eps)
        bg = torch.ones_like(rgb) if self.white_bg else torch.zeros_like(rgb)
        rgb = rgb * occu[:, None] + bg * (1. - occu[:, None])          # imgutil.alpha_blend
        return rgb, occu, depth, disp, weights

    def _render_rays(self, rayo, rayd, mode='train', precision=None, perturb_u=None):
        """nerf.py:149-182.  Stratified perturbation only with explicit uniforms (`perturb_u`
        [n, n_samples_coarse]); the fine pass is deterministic (gen_z_fine)."""
        n_c = self.config.getint('DEFAULT', 'n_samples_coarse')
        lin = self.config.getboolean('DEFAULT', 'lin_in_disp')
        rayo = rayo.to(self.device, torch.float32).contiguous()
        rayd = rayd.to(self.device, torch.float32)
        rayd = (rayd * torch.rsqrt(torch.clamp((rayd * rayd).sum(1, keepdim=True), min=1e-12))
                ).contiguous()                                         # tf.linalg.l2_normalize
        n = rayo.shape[0]
        z = _lib.gen_z(self.ctx, self.near, self.far, n_c, n, lin,
                       perturb_u if mode == 'train' else None)
        rgbs = self._eval_nerf_at(rayo, rayd, z, False, precision)
        rgb, occu, depth, disp, weights = self._accumulate(rgbs, z, rayd)
        pred_coarse = {'rgb': rgb, 'occu': occu, 'depth': depth, 'disp': disp}
        if self.n_samples_fine <= 0:
            return pred_coarse, {}
        z = _lib.gen_z_fine(self.ctx, z, weights, self.n_samples_fine)
        rgbs = self._eval_nerf_at(rayo, rayd, z, True, precision)
        rgb, occu, depth, disp, _ = self._accumulate(rgbs, z, rayd)
        return pred_coarse, {'rgb': rgb, 'occu': occu, 'depth': depth, 'disp': disp}

    def call(self, batch, mode='train', precision=None):
        """nerf.py:100-118: batch = (id_, hw, rayo, rayd, rgb), all flattened."""
        self._validate_mode(mode)
        id_, hw, rayo, rayd, rgb = batch
        to_t = lambda x: x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, np.float32))
        pred_coarse, pred_fine = self._render_rays(to_t(rayo), to_t(rayd), mode, precision)
        pred = {'coarse': pred_coarse['rgb'], 'fine': pred_fine.get('rgb', None)}
        gt = rgb
        to_vis = {'id': id_, 'hw': hw, 'gt_rgb': gt}
        for k, v in pred_coarse.items():
            to_vis['coarse_' + k] = v
        for k, v in pred_fine.items():
            to_vis['fine_' + k] = v
        return pred, gt, {}, to_vis

    def compute_loss(self, pred, gt, **kwargs):
        """nerf.py:292-300 with `loss = l2` (losses.py:32-47): per-ray MSE of the coarse and the
        fine rendering when `keep_batch` is set, their scalar means otherwise."""
        if self.config.get('DEFAULT', 'loss', fallback='l2') != 'l2':
            raise NotImplementedError("only the l2 loss of the shipped nerf.ini")
        keep_batch = kwargs.get('keep_batch', False)
        gt = gt if torch.is_tensor(gt) else torch.as_tensor(np.asarray(gt, np.float32))
        gt = gt.to(self.device)
        loss = 0
        for p in (pred['coarse'], pred['fine']):
            if p is None:
                continue
            per_ray = torch.mean((p - gt) ** 2, dim=-1)
            loss = loss + (per_ray if keep_batch else per_ray.mean())
        return loss

    def weights_changed(self):
        """Call after editing layer weights so they are re-packed for the GPU."""
        self._packed.clear()

    # ---- static samplers, same signatures as the reference ------------------
    @staticmethod
    def gen_z(near, far, n_samples, n_rays, lin_in_disp=False, perturb=False,
              perturb_u=None):
        """nerf.py:120-136.  perturb=True draws the uniforms with torch (or takes
        them from `perturb_u` [n_rays, n_samples])."""
        ctx = _lib.default_context()
        if perturb and perturb_u is None:
            perturb_u = torch.rand((n_rays, n_samples), device=ctx.device)
        return _lib.gen_z(ctx, near, far, n_samples, n_rays, lin_in_disp,
                          perturb_u if perturb else None)

    @staticmethod
    def gen_z_fine(z_coarse, weights, n_samples_fine, perturb=False):
        """nerf.py:138-147 (deterministic inverse-CDF sampling only)."""
        if perturb:
            raise NotImplementedError("perturb=True (random u) in gen_z_fine")
        return _lib.gen_z_fine(_lib.default_context(), z_coarse, weights, n_samples_fine)

    @staticmethod
    def accumulate_sigma(sigma, z, rayd, noise_std=0., inf=1e10, accu_chunk=65536):
        """nerf.py:184-212."""
        if noise_std != 0. or inf != 1e10:
            raise NotImplementedError("noise_std != 0 / inf != 1e10")
        w, _, _, _, _ = _lib.composite(_lib.default_context(), sigma, z, rayd, rayd,
                                       want_weights=True, want_surf=False)
        return w

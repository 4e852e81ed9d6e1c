"""Mirror of the parts of nerfactor/models/nerf.py Stage A uses: the coarse / fine
sigma networks (`net['coarse_enc']`, `net['coarse_sigma_out']`, ... nerf.py:53-71),
the `xyz` embedder and the static samplers `gen_z`, `gen_z_fine`,
`accumulate_sigma` (nerf.py:120-147, 184-212).  NeRF's RGB branch and its own
training are a "next" row."""
import numpy as np
import torch

from .. import _lib
from ..networks import mlp
from ..networks.embedder import Embedder
from .base import Model as BaseModel


class Model(BaseModel):
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
        self._packed = {}
        if params is not None:
            self.load_params(params)

    def _init_net(self):
        """nerf.py:53-71 (sigma branch; bottleneck / rgb_out belong to the RGB branch)."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'enc_depth')
        act = self.config.get('DEFAULT', 'act', fallback='relu')
        if act != 'relu':
            raise NotImplementedError(act)
        return {'enc': mlp.Network([w] * d, act=[act] * d, skip_at=[d // 2]),
                'sigma_out': mlp.Network([1], act=[None])}

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

"""Mirror of nerfactor/models/nerfactor_microfacet.py: analytic GGX BRDF with a
predicted scalar roughness (z_dim = 1, sigmoid head) instead of the learned latent."""
from ..brdf.microfacet import Microfacet
from ..networks import mlp
from .nerfactor import Model as NeRFactorModel


class Model(NeRFactorModel):
    _uses_brdf_prior = False                # nerfactor_microfacet.py:35-40: no brdf_model_ckpt read

    def _init_brdf_dims(self):
        self.z_dim = 1                      # nerfactor_microfacet.py:38
        self.normalize_brdf_z = False

    def _init_brdf_model(self, params):
        self.brdf_model = None              # no learned prior in this variant

    def _init_embedder(self):
        """nerfactor_microfacet.py:102-106: grandparent's embedders (no rusink)."""
        return super(NeRFactorModel, self)._init_embedder()

    def _init_net(self):
        """nerfactor_microfacet.py:108-114: roughness head gets a sigmoid."""
        net = super()._init_net()
        net['brdf_z_out'] = mlp.Network([self.z_dim], act=['sigmoid'])
        return net

    def _eval_brdf_at(self, pts2l, pts2c, normal, albedo, brdf_prop, pts=None, cam=None):
        """nerfactor_microfacet.py:116-124: the GGX lobe is evaluated inside the
        rendering-equation kernel; hand it the roughness and f0."""
        fresnel_f0 = self.config.getfloat('DEFAULT', 'fresnel_f0')
        return {'microfacet': Microfacet(f0=fresnel_f0), 'rough': brdf_prop.contiguous()}

    def _brdf_prop_as_img(self, brdf_prop):
        """nerfactor_microfacet.py:126-132: roughness as a grey image."""
        import numpy as np
        return np.concatenate([brdf_prop] * 3, axis=2)

    def _brdf_kernel_args(self, brdf):
        return {'rough': brdf['rough'], 'f0': brdf['microfacet'].f0}

    def _fused_stage_b(self, pts, normal, cam, albedo, brdf_prop, light, want_lvis, all_lights=False):
        """GGX lobe: with one env-map and <= 512 lights this is ONE kernel (the rendering equation
        runs in the head epilogue of the light-visibility network)."""
        from .. import _lib
        m_lvis = self._packed_mlp('lvis', 'lvis', n_freqs_a=self.embedder['xyz'].n_freqs,
                                  n_freqs_b=self.embedder['ldir'].n_freqs)
        return _lib.stageB_fused_fwd(
            self.ctx, m_lvis, pts, normal, cam, albedo, self.lxyz, self.lareas, light,
            rough=brdf_prop, light_idx=self.light_idx,
            f0=self.config.getfloat('DEFAULT', 'fresnel_f0'), xyz_scale=self.xyz_scale,
            linear2srgb=self.config.getboolean('DEFAULT', 'linear2srgb'),
            precision=self.precision, want_lvis=want_lvis, all_lights=all_lights)

"""Mirror of nerfactor/models/brdf.py: the MERL-prior BRDF MLP
(z ++ posenc(rusink) -> softplus scalar, brdf.py:57-66) and its latent codes.
The hot path only needs the frozen forward (nerfactor.py:436-452); the prior's own
training / visualisation is out of scope."""
import numpy as np

from ..networks import mlp
from ..networks.embedder import Embedder
from ..networks.layers import LatentCode
from .base import Model as BaseModel


class Model(BaseModel):
    def __init__(self, config, debug=False, params=None, brdf_names=None):
        super().__init__(config, debug=debug)
        self.mlp_chunk = self.config.getint('DEFAULT', 'mlp_chunk')
        self.embedder = self._init_embedder()
        self.net = self._init_net()
        self.brdf_names = list(brdf_names or ['synthetic_%03d' % i for i in range(4)])
        z_dim = self.config.getint('DEFAULT', 'z_dim')
        self.latent_code = LatentCode(
            len(self.brdf_names), z_dim,
            mean=self.config.getfloat('DEFAULT', 'z_gauss_mean'),
            std=self.config.getfloat('DEFAULT', 'z_gauss_std'),
            normalize=self.config.getboolean('DEFAULT', 'normalize_z'),
            rng=np.random.default_rng(0))
        width = self.config.getint('DEFAULT', 'mlp_width')
        rng = np.random.default_rng(1)
        self.net['brdf_mlp'].build(z_dim + self.embedder['rusink'].out_dims, rng)
        self.net['brdf_out'].build(width, rng)
        if params is not None:
            for k in ('brdf_mlp', 'brdf_out'):
                self.net[k].load(params[k])
        self.trainable = False

    def _init_net(self):
        """brdf.py:57-66."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'mlp_depth')
        s = self.config.getint('DEFAULT', 'mlp_skip_at')
        return {'brdf_mlp': mlp.Network([w] * d, act=['relu'] * d, skip_at=[s]),
                'brdf_out': mlp.Network([1], act=['softplus'])}

    def _init_embedder(self):
        """brdf.py:68-85."""
        n = self.config.getint('DEFAULT', 'n_freqs')
        return {'rusink': Embedder(incl_input=True, in_dims=3, log2_max_freq=n - 1,
                                   n_freqs=n, log_sampling=True)}

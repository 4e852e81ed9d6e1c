"""Mirror of nerfactor/networks/layers.py:24-67 (LatentCode only; the conv /
norm helpers at :70-169 have no caller in the reference)."""
import numpy as np


class LatentCode:
    def __init__(self, n_iden, dim, mean=0., std=1., normalize=False, rng=None):
        rng = rng or np.random.default_rng()
        self._z = (mean + std * rng.standard_normal((n_iden, dim))).astype(np.float32)
        self.normalize = normalize

    @property
    def z(self):
        if self.normalize:
            sq = np.sum(self._z ** 2, axis=1, keepdims=True)
            return self._z / np.sqrt(np.maximum(sq, 1e-6))
        return self._z

    @z.setter
    def z(self, value):
        self._z = np.asarray(value, dtype=np.float32)

    def __call__(self, ind):
        ind = np.atleast_1d(np.asarray(ind))
        return self.z[ind]

    def interp(self, w1, i1, w2, i2):
        if self.normalize:
            raise NotImplementedError("slerp of normalised codes (geom.py:82-116)")
        return w1 * self(i1) + w2 * self(i2)

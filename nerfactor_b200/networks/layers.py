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
        """layers.py:58-67: linear interpolation, or -- codes on the unit sphere -- slerp
        (util/geom.py:100-116; safe_acos clips its argument to [-1, 1], util/math.py:40-58)."""
        z1, z2 = self(i1), self(i2)
        if not self.normalize:
            return w1 * z1 + w2 * z2
        assert w1 + w2 == 1., "When latent codes are normalized, use weights that sum to 1"
        omega = np.arccos(np.clip(np.sum(z1 * z2), -1., 1.))
        if omega == 0.:                       # identical codes: the reference divides 0 / 0 here
            return z1
        return ((z1 * np.sin((1. - w2) * omega) + z2 * np.sin(w2 * omega)) /
                np.sin(omega)).astype(np.float32)

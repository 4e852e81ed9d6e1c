"""Mirror of nerfactor/datasets/brdf_merl.py: the training data of the BRDF prior.  One .npz per
material (`train_<name>.npz`, `vali_<name>.npz`: `name`, `i`, `envmap_h`, `ims`, `spp`,
`rusink [R,3]`, `refl [R,1]`) plus a single `test*.npz` with the Rusinkiewicz coordinates every
test identity shares; the test "paths" are material names followed by interpolation recipes
`<k>_<w1>_<mat1>_<w2>_<mat2>` (brdf_merl.py:44-63).  Element:
`(id_, i, envmap_h, ims, spp, rusink, refl)` with per-material scalars for the first five."""
from os.path import basename

import numpy as np

from ..util import io as ioutil
from .base import Dataset as BaseDataset


class Dataset(BaseDataset):
    def __init__(self, config, mode, debug=False, seed=0, n_iden=20, n_between=11, **kw):
        root = config.get('DEFAULT', 'data_root')
        train_paths = ioutil.sortglob(root, 'train_*', ext='npz')
        vali_paths = ioutil.sortglob(root, 'vali_*', ext='npz')
        test_paths = ioutil.sortglob(root, 'test*', ext='npz')
        assert len(test_paths) == 1, (
            "There should be a single set of test coordinates, shared by all identities")
        self.brdf_names = [basename(x)[len('train_'):-len('.npz')] for x in train_paths]
        self.test_data = ioutil.load_np(test_paths[0])
        ids = list(self.brdf_names)                    # novel Rusink., seen identities
        rng = np.random.RandomState(seed)              # np.random.seed(seed) + np.random.choice
        mats = rng.choice(self.brdf_names, min(n_iden, len(self.brdf_names)), replace=False)
        k = 0
        for m in range(len(mats) - 1):                 # novel Rusink., interpolated identities
            for a in np.linspace(1, 0, n_between, endpoint=True):
                ids.append(f'{k:06d}_{a:f}_{mats[m]}_{1 - a:f}_{mats[m + 1]}')
                k += 1
        self.paths = {'train': train_paths, 'vali': vali_paths, 'test': ids}
        super().__init__(config, mode, debug=debug, seed=seed, **kw)

    def _get_batch_size(self):
        return self.config.getint('DEFAULT', 'n_rays_per_step')

    def get_n_brdfs(self):
        return len(self.paths[self.mode])

    def _glob(self):
        return self.paths[self.mode]

    def _process_example_precache(self, path):
        return self._load_data(path)

    def _load_data(self, path):
        """brdf_merl.py:89-108."""
        data = self.test_data if self.mode == 'test' else ioutil.load_np(path)
        rusink = np.asarray(data['rusink'])
        if self.mode == 'test':
            id_ = path
            i = self.brdf_names.index(id_) if id_ in self.brdf_names else -1
            refl = np.zeros((rusink.shape[0], 1), dtype=rusink.dtype)      # placeholder
        else:
            id_, i, refl = str(data['name'][()]), int(data['i'][()]), np.asarray(data['refl'])
        return (id_, i, int(data['envmap_h'][()]), int(data['ims'][()]), int(data['spp'][()]),
                rusink.astype(np.float32), refl.astype(np.float32))

    def _process_example_postcache(self, id_, i, envmap_h, ims, spp, rusink, refl):
        """brdf_merl.py:111-148: all entries (vali / test) or `bs` draws with replacement."""
        if self.mode == 'train':
            sel = self.rng.integers(0, rusink.shape[0], size=self.bs)
            rusink, refl = rusink[sel], refl[sel]
        return id_, i, envmap_h, ims, spp, rusink, refl

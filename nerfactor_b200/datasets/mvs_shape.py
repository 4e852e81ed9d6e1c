"""Mirror of nerfactor/datasets/mvs_shape.py: Stage-B inputs whose geometry comes from multi-view
stereo instead of a NeRF (the reference's real-scene path, config/nerfactor_mvs.ini).  Everything
of a view -- `metadata.json` (with `cam_loc`), `rgba.png`, `alpha.png`, `xyz.npy`, `normal.npy`,
`lvis.npy` -- sits in `<mvs_root>/<view>/`; the light positions come from `<mvs_root>/lights.npz`
(models/shape.py:67-74)."""
from os.path import dirname, join

from ..util import geom_io, io as ioutil
from .nerf_shape import Dataset as NerfShapeDataset


class Dataset(NerfShapeDataset):
    def _glob(self):
        """mvs_shape.py:30-63."""
        mvs_root = self.config.get('DEFAULT', 'mvs_root')
        pattern = ('%s_000' if self.debug else '%s_???') % self._mode_str()
        keep = []
        for metadata_path in ioutil.sortglob(join(mvs_root, pattern), 'metadata.json'):
            view_dir = join(mvs_root, self._parse_id(metadata_path))
            paths = {k: join(view_dir, f) for k, f in (
                ('xyz', 'xyz.npy'), ('normal', 'normal.npy'), ('lvis', 'lvis.npy'),
                ('alpha', 'alpha.png'))}
            if self.mode != 'test':
                paths['rgba'] = join(view_dir, 'rgba.png')
            if ioutil.all_exist(paths):
                keep.append(metadata_path)
                self.meta2buf[metadata_path] = paths
        return keep

    def _load_data(self, metadata_path):
        """mvs_shape.py:66-121."""
        paths = self.meta2buf[metadata_path]
        return geom_io.load_view(
            metadata_path, dirname(paths['xyz']), self.config.getint('DEFAULT', 'imh'),
            mode=self.mode, rgba_path=paths.get('rgba'),
            use_nerf_alpha=self.config.getboolean('DEFAULT', 'use_nerf_alpha', fallback=False),
            debug=self.debug, rays_from_cam_loc=True)

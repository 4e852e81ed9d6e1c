"""Mirror of nerfactor/datasets/nerf_shape.py.

`Dataset` (:29-190): views of `<data_root>` that have all four Stage-A buffers under
`<data_nerf_root>/<view>/` -> the 9-tuple `(id_, hw, rayo, rayd, rgb, alpha, xyz, normal, lvis)`
`Model.call` consumes.  `sample_rays` / `make_batch` (:72-121) are the same steps on torch
tensors on any device (used when Stage A hands its buffers to Stage B without touching disk)."""
from os.path import dirname, join

import numpy as np
import torch

from ..util import geom_io, io as ioutil
from .nerf import Dataset as NerfDataset


def sample_rays(rayo, rayd, rgb, alpha, xyz, normal, lvis, mode='train', bs=1024,
                alpha_thres=0.9, always_all_rays=False, generator=None):
    """nerf_shape.py:84-121.  vali / test (or always_all_rays): every ray, row-major
    (ray n = y * W + x).  train: `bs` uniform draws WITH replacement from the pixels whose
    alpha exceeds `alpha_thres` (all pixels if alpha_thres is None)."""
    h, w = rgb.shape[0], rgb.shape[1]
    if mode in ('vali', 'test') or always_all_rays:
        f = lambda t, c: t.reshape(-1, c)
        return (f(rayo, 3), f(rayd, 3), f(rgb, 3), f(alpha, 1), f(xyz, 3), f(normal, 3),
                lvis.reshape(h * w, -1))
    if alpha_thres is None:
        coords = torch.arange(h * w, device=rgb.device)
    else:
        coords = torch.nonzero(alpha.reshape(-1) > alpha_thres, as_tuple=False)[:, 0]
    if coords.numel() == 0:
        raise ValueError("no foreground pixel above alpha_thres in this view")
    sel = coords[torch.randint(0, coords.numel(), (bs,), device=coords.device,
                               generator=generator)]
    g = lambda t, c: t.reshape(h * w, c).index_select(0, sel)
    return (g(rayo, 3), g(rayd, 3), g(rgb, 3), g(alpha, 1), g(xyz, 3), g(normal, 3),
            lvis.reshape(h * w, -1).index_select(0, sel))


def make_batch(id_, hw, rays):
    """nerf_shape.py:72-82: (id_, hw, rayo, rayd, rgb, alpha, xyz, normal, lvis); id_ / hw are
    per-view scalars here (the reference tiles them per ray only for tf.distribute)."""
    return (id_, hw) + tuple(rays)


class Dataset(NerfDataset):
    def __init__(self, config, mode, debug=False, always_all_rays=False, **kw):
        self.meta2buf = {}
        super().__init__(config, mode, debug=debug, always_all_rays=always_all_rays, **kw)

    def _glob(self):
        """nerf_shape.py:35-69: only cameras with every required buffer."""
        root = self.config.get('DEFAULT', 'data_root')
        nerf_root = self.config.get('DEFAULT', 'data_nerf_root')
        mode_str = self._mode_str()
        metadata_dir = join(root, ('%s_002' if self.debug else '%s_???') % mode_str)
        keep = []
        for metadata_path in ioutil.sortglob(metadata_dir, 'metadata.json'):
            id_ = self._parse_id(metadata_path)
            paths = {k: join(nerf_root, id_, f) for k, f in (
                ('xyz', 'xyz.npy'), ('normal', 'normal.npy'), ('lvis', 'lvis.npy'),
                ('alpha', 'alpha.png'))}
            if self.mode != 'test':
                paths['rgba'] = join(dirname(metadata_path), 'rgba.png')
            if ioutil.all_exist(paths):
                keep.append(metadata_path)
                self.meta2buf[metadata_path] = paths
        return keep

    def _process_example_postcache(self, id_, rayo, rayd, rgb, alpha, xyz, normal, lvis):
        """nerf_shape.py:72-82."""
        hw = tuple(int(x) for x in rgb.shape[:2])
        rays = self._sample_rays(rayo, rayd, rgb, alpha, xyz, normal, lvis)
        return (id_, hw) + tuple(rays)

    def _sample_rays(self, rayo, rayd, rgb, alpha, xyz, normal, lvis, alpha_thres=0.9):
        """nerf_shape.py:84-121 on host arrays (numpy draws instead of tf.random.uniform)."""
        h, w = rgb.shape[:2]
        f = lambda a, c: np.ascontiguousarray(a.reshape(h * w, c))
        flat = (f(rayo, 3), f(rayd, 3), f(rgb, 3), f(alpha, 1), f(xyz, 3), f(normal, 3),
                f(lvis, lvis.shape[2]))
        if self.mode in ('vali', 'test') or self.always_all_rays:
            return flat
        if alpha_thres is None:
            coords = np.arange(h * w)
        else:
            coords = np.nonzero(alpha.reshape(-1) > alpha_thres)[0]
        if coords.size == 0:
            raise ValueError("no foreground pixel above alpha_thres in this view")
        sel = coords[self.rng.integers(0, coords.size, size=self.bs)]
        return tuple(a[sel] for a in flat)

    def _load_data(self, metadata_path):
        """nerf_shape.py:133-190 (shared with the standalone loader util/geom_io.load_view)."""
        paths = self.meta2buf[metadata_path]
        return geom_io.load_view(
            metadata_path, dirname(paths['xyz']), self.config.getint('DEFAULT', 'imh'),
            mode=self.mode, rgba_path=paths.get('rgba'),
            use_nerf_alpha=self.config.getboolean('DEFAULT', 'use_nerf_alpha', fallback=False),
            debug=self.debug)

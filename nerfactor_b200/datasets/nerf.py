"""Mirror of nerfactor/datasets/nerf.py:29-215: one view = `<data_root>/{train,val,test}_???/
metadata.json` (+ `rgba.png`), rays from the camera in the metadata, RGBA composited onto the
background colour.  Element: `(id_, hw, rayo, rayd, rgb)`; id_ / hw are per-view values (the
reference tiles them per ray only to satisfy tf.distribute, nerf.py:112-115)."""
from os.path import basename, dirname, exists, join

import numpy as np

from ..util import img as imgutil, io as ioutil
from .base import Dataset as BaseDataset


class Dataset(BaseDataset):
    def __init__(self, config, mode, debug=False, always_all_rays=False, spp=1, **kw):
        self.meta2img = {}
        sps = np.sqrt(spp)                      # samples per side (nerf.py:32-37)
        assert sps == int(sps), (
            "Samples per pixel must be a square number so that samples per side are integers")
        self.sps = int(sps)
        self.always_all_rays = always_all_rays
        super().__init__(config, mode, debug=debug, **kw)

    def get_n_views(self):
        return len(self.files)

    def _get_batch_size(self):
        """nerf.py:52-62."""
        if self.mode == 'train':
            return self.config.getint('DEFAULT', 'n_rays_per_step')
        ret = self._load_data(self.files[0])
        return int(np.prod(ret[-1].shape[:2]))

    def _mode_str(self):
        return 'val' if self.mode == 'vali' else self.mode

    def _glob(self):
        """nerf.py:64-90."""
        root = self.config.get('DEFAULT', 'data_root')
        metadata_dir = join(root, '%s_???' % self._mode_str())
        paths = ioutil.sortglob(metadata_dir, 'metadata.json')
        if self.mode == 'test':
            return paths
        keep = []
        for metadata_path in paths:           # only cameras with a paired RGBA image
            img_path = join(dirname(metadata_path), 'rgba.png')
            if exists(img_path):
                keep.append(metadata_path)
                self.meta2img[metadata_path] = img_path
        return keep

    @staticmethod
    def _parse_id(metadata_path):
        return basename(dirname(metadata_path))

    def _process_example_precache(self, path):
        return self._load_data(path)

    def _process_example_postcache(self, id_, rayo, rayd, rgb):
        """nerf.py:103-116."""
        hw = tuple(int(x) for x in rgb.shape[:2])
        rayo, rayd, rgb = self._sample_rays(rayo, rayd, rgb)
        return id_, hw, rayo, rayd, rgb

    def _sample_rays(self, rayo, rayd, rgb):
        """nerf.py:118-141: all rays row-major, or `bs` uniform draws with replacement."""
        f = lambda a: np.ascontiguousarray(a.reshape(-1, 3))
        if self.mode in ('vali', 'test') or self.always_all_rays:
            return f(rayo), f(rayd), f(rgb)
        sel = self.rng.integers(0, rgb.shape[0] * rgb.shape[1], size=self.bs)
        return f(rayo)[sel], f(rayd)[sel], f(rgb)[sel]

    def _read_camera(self, metadata_path):
        imh = self.config.getint('DEFAULT', 'imh')
        metadata = ioutil.read_json(metadata_path)
        imw = int(imh / metadata['imh'] * metadata['imw'])
        cam_to_world = np.array(
            [float(x) for x in metadata['cam_transform_mat'].split(',')]).reshape(4, 4)
        return imh, imw, cam_to_world, metadata['cam_angle_x']

    def _load_data(self, metadata_path):
        """nerf.py:150-170."""
        white_bg = self.config.getboolean('DEFAULT', 'white_bg')
        id_ = self._parse_id(metadata_path)
        imh, imw, cam_to_world, cam_angle_x = self._read_camera(metadata_path)
        rayo, rayd = self._gen_rays(cam_to_world, cam_angle_x, imh, imw)
        rayo, rayd = rayo.astype(np.float32), rayd.astype(np.float32)
        if self.mode == 'test':
            return id_, rayo, rayd, np.zeros((imh, imw, 3), dtype=np.float32)
        rgba = imgutil.read(self.meta2img[metadata_path])
        assert rgba.ndim == 3 and rgba.shape[2] == 4, "Input image is not RGBA"
        rgba = imgutil.normalize_uint(rgba)
        if imh != rgba.shape[0]:
            rgba = imgutil.resize_cv2(rgba, new_h=imh)
        rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
        bg = np.ones_like(rgb) if white_bg else np.zeros_like(rgb)
        rgb = imgutil.alpha_blend(rgb, alpha, tensor2=bg)
        return id_, rayo, rayd, rgb.astype(np.float32)

    def _gen_rays(self, to_world, angle_x, imh, imw):
        """nerf.py:172-214 in fp64 like the reference (the device version, nf_gen_rays, is
        bit-identical for spp = 1: tests/test_gpu_parity.py).  Pixel corners, no half-pixel
        offset; un-normalised directions; `sps` sub-samples per pixel side."""
        n_y, n_x = imh * self.sps, imw * self.sps
        px, py = np.meshgrid(np.linspace(0, imw, n_x, endpoint=False),
                             np.linspace(0, imh, n_y, endpoint=False))
        focal = .5 * imw / np.tan(.5 * angle_x)
        local = np.stack(((px - .5 * imw) / focal, -(py - .5 * imh) / focal, -np.ones_like(px)),
                         axis=-1)
        rayd = np.sum(local[:, :, np.newaxis, :] * to_world[:3, :3], axis=-1)     # R . d_local
        rayo = np.tile(to_world[:3, 3][None, None, :], (n_y, n_x, 1))
        if self.config.getboolean('DEFAULT', 'ndc'):
            rayo, rayd = _to_ndc(rayo, rayd, self.config.getfloat('DEFAULT', 'near'), focal,
                                 imh, imw)
        return rayo, rayd


def _to_ndc(rayo, rayd, near, focal, imh, imw):
    """NeRF's normalised-device-coordinate rays (nerf.py:194-213; marked "not in use" upstream,
    ndc = False in every shipped config).  OpenGL convention: flip y / z of SfM cameras, move the
    origins onto the near plane, then project origins and directions."""
    flip = np.diag((1.0, -1.0, -1.0))
    o, d = rayo.dot(flip), rayd.dot(flip)
    o = o + (-(near + o[..., 2]) / d[..., 2])[..., None] * d
    sx, sy = -1. / (imw / (2. * focal)), -1. / (imh / (2. * focal))
    ox_z, oy_z = o[..., 0] / o[..., 2], o[..., 1] / o[..., 2]
    o_ndc = np.dstack((sx * ox_z, sy * oy_z, 1. + 2. * near / o[..., 2]))
    d_ndc = np.dstack((sx * (d[..., 0] / d[..., 2] - ox_z), sy * (d[..., 1] / d[..., 2] - oy_z),
                       -2. * near / o[..., 2]))
    return o_ndc, d_ndc

"""On-disk geometry buffers between Stage A and Stage B.

Writer side = nerfactor/util/geom.py:27-79 (`write_alpha`, `write_xyz`, `write_normal`,
`write_lvis`): per view directory `alpha.png` (uint8), `xyz.npy`, `normal.npy`, `lvis.npy`
(float32, raw) plus the PNG visualisations `xyz.png`, `normal.png`, `lvis.png` (the per-light
`lvis.mp4` of geom.py:37-44 is visualisation only and not produced).

Reader side = `Dataset._load_data`, nerfactor/datasets/nerf_shape.py:133-190: rays from the
view's `metadata.json`, the four buffers, optional RGBA, resize to `imh`, normals re-normalised,
visibility clipped.  Everything here is host-side numpy; `to_pinned` hands the result to the
GPU pipeline through page-locked memory.
"""
import json
import os

import numpy as np


def _write_png(arr_0to1, path, clip=False):
    """xiuminglib io.img.write_float: (x * 255).astype(uint8) -- truncation, not rounding."""
    from PIL import Image
    a = np.asarray(arr_0to1, np.float64)
    if clip:
        a = np.clip(a, 0., 1.)
    elif a.size and (a.min() < 0 or a.max() > 1):
        raise ValueError("Input should be in [0, 1], or allow it to be clipped")
    img = (a * 255).astype(np.uint8)
    if img.ndim == 3 and img.shape[2] == 1:
        img = np.dstack([img] * 3)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    Image.fromarray(img).save(path)
    return img


def write_alpha(alpha, out_dir):
    """geom.py:75-79."""
    return _write_png(np.asarray(alpha, np.float32), os.path.join(out_dir, 'alpha.png'))


def write_xyz(xyz, out_dir):
    """geom.py:47-59."""
    arr = np.asarray(xyz, np.float32)
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, 'xyz.npy'), arr)
    rng = arr.max() - arr.min()
    _write_png((arr - arr.min()) / (rng if rng > 0 else 1.), os.path.join(out_dir, 'xyz.png'),
               clip=True)


def write_normal(normal, out_dir):
    """geom.py:62-72."""
    arr = np.asarray(normal, np.float32)
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, 'normal.npy'), arr)
    _write_png((arr + 1) / 2, os.path.join(out_dir, 'normal.png'))


def write_lvis(lvis, out_dir):
    """geom.py:27-36 (raw dump + the light-averaged image)."""
    arr = np.asarray(lvis, np.float32)
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, 'lvis.npy'), arr)
    _write_png(arr.mean(axis=2), os.path.join(out_dir, 'lvis.png'))


EXPECTED = ('alpha.png', 'lvis.npy', 'lvis.png', 'normal.npy', 'normal.png', 'xyz.npy', 'xyz.png')


def view_done(out_dir):
    """geometry_from_nerf.py:106-115: has this view been processed already?"""
    return all(os.path.exists(os.path.join(out_dir, f)) for f in EXPECTED)


def write_view_buffers(buffers, out_dir):
    """Writes what geometry_from_nerf.process_view leaves on disk (:127-174) from the dict
    `nerfactor_b200.geometry_from_nerf.process_view` returns (device tensors or arrays)."""
    host = {k: (v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v))
            for k, v in buffers.items()}
    write_alpha(host['alpha'], out_dir)
    write_xyz(host['xyz'], out_dir)
    write_normal(host['normal'], out_dir)
    if 'lvis' in host:
        write_lvis(host['lvis'], out_dir)


# ------------------------------------------------------------------------------- reader
def _read_img(path):
    from PIL import Image
    return np.array(Image.open(path))


def _normalize_uint(arr):
    """xiuminglib img.normalize_uint."""
    if arr.dtype not in (np.uint8, np.uint16):
        raise TypeError(arr.dtype)
    return arr.astype(float) / np.iinfo(arr.dtype).max


def _resize(arr, new_h):
    """xiuminglib img.resize(method='cv2'): aspect preserved, INTER_AREA when shrinking,
    INTER_LINEAR when enlarging.  cv2 handles at most 512 channels per call."""
    import cv2
    h, w = arr.shape[:2]
    new_w = int(w / h * new_h)
    interp = cv2.INTER_LINEAR if new_h > h else cv2.INTER_AREA
    if arr.ndim == 3 and arr.shape[2] > 512:
        parts = [cv2.resize(arr[:, :, i:i + 512], (new_w, new_h), interpolation=interp)
                 for i in range(0, arr.shape[2], 512)]
        parts = [p[:, :, None] if p.ndim == 2 else p for p in parts]
        return np.concatenate(parts, axis=2)
    return cv2.resize(arr, (new_w, new_h), interpolation=interp)


def gen_rays_host(cam_to_world, cam_angle_x, imh, imw):
    """Dataset._gen_rays, datasets/nerf.py:172-193 (ndc=False, spp=1), fp64 like the reference;
    the device version is nf_gen_rays (bit-identical, tests/test_gpu_parity.py)."""
    fl = .5 * imw / np.tan(.5 * cam_angle_x)
    xs = np.linspace(0, imw, imw, endpoint=False)
    ys = np.linspace(0, imh, imh, endpoint=False)
    xs, ys = np.meshgrid(xs, ys)
    rayd = np.stack(((xs - .5 * imw) / fl, -(ys - .5 * imh) / fl, -np.ones_like(xs)), axis=-1)
    rayd = np.sum(rayd[:, :, np.newaxis, :] * cam_to_world[:3, :3], axis=-1)
    rayo = np.tile(cam_to_world[:3, 3], (rayd.shape[0], rayd.shape[1], 1))
    return rayo, rayd


def load_view(metadata_path, buffer_dir, imh, mode='test', rgba_path=None, use_nerf_alpha=False,
              debug=False, n_lights_debug=512, rays_from_cam_loc=False):
    """nerf_shape.py:133-190.  Returns (id_, rayo, rayd, rgb, alpha, xyz, normal, lvis), arrays
    [H, W, ...] float32 (alpha [H, W]).  `mode` 'test': NeRF-traced alpha, zero RGB;
    'train' / 'vali': RGBA image at `rgba_path` (ground-truth alpha unless use_nerf_alpha)."""
    id_ = os.path.basename(os.path.dirname(metadata_path))
    with open(metadata_path) as f:
        metadata = json.load(f)
    if rays_from_cam_loc:
        # MVS geometry (datasets/mvs_shape.py:72-78): only the camera location is known; rays keep
        # the metadata's own size and the directions are dummies (Stage B never reads them)
        rayo = np.tile(np.array(metadata['cam_loc'])[None, None, :],
                       (metadata['imh'], metadata['imw'], 1)).astype(np.float32)
        rayd = np.zeros_like(rayo)
    else:
        imw = int(imh / metadata['imh'] * metadata['imw'])
        cam_to_world = np.array(
            [float(x) for x in metadata['cam_transform_mat'].split(',')]).reshape(4, 4)
        rayo, rayd = gen_rays_host(cam_to_world, metadata['cam_angle_x'], imh, imw)
        rayo, rayd = rayo.astype(np.float32), rayd.astype(np.float32)
    xyz = np.load(os.path.join(buffer_dir, 'xyz.npy'))
    normal = np.load(os.path.join(buffer_dir, 'normal.npy'))
    if debug:
        lvis = 0.5 * np.ones(normal.shape[:2] + (n_lights_debug,), dtype=np.float32)
    else:
        lvis = np.load(os.path.join(buffer_dir, 'lvis.npy'))
    if mode == 'test':
        alpha = _normalize_uint(_read_img(os.path.join(buffer_dir, 'alpha.png')))
        rgb = np.zeros_like(xyz)
    else:
        rgba = _read_img(rgba_path)
        assert rgba.ndim == 3 and rgba.shape[2] == 4, "Input image is not RGBA"
        rgba = _normalize_uint(rgba)
        rgb = rgba[:, :, :3]
        if use_nerf_alpha:
            alpha = _normalize_uint(_read_img(os.path.join(buffer_dir, 'alpha.png')))
        else:
            alpha = rgba[:, :, 3]
    if alpha.ndim == 3:                       # alpha.png is written as 3 identical channels
        alpha = alpha[:, :, 0]
    if imh != xyz.shape[0]:
        xyz, normal, lvis, alpha, rgb = [_resize(np.ascontiguousarray(a), imh)
                                         for a in (xyz, normal, lvis, alpha, rgb)]
    assert not np.isclose(xyz, rayo).all(axis=2).any(), "Found XYZs coinciding with the camera"
    norm = np.linalg.norm(normal, axis=2, keepdims=True)
    normal = normal / np.where(norm == 0, 1., norm)                 # xm.linalg.normalize
    assert np.isclose(np.linalg.norm(normal, axis=2), 1).all(), \
        "Found normals with a norm far away from 1"
    lvis = np.clip(lvis, 0, 1)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return id_, rayo, rayd, f32(rgb), f32(alpha), f32(xyz), f32(normal), f32(lvis)


def to_pinned(arrays):
    """Host arrays -> page-locked torch tensors (async H2D copies, bench.py `e2e` path)."""
    import torch
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a))
        out.append(t.pin_memory() if torch.cuda.is_available() else t)
    return out

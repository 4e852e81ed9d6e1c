"""Geometry-buffer files between Stage A and Stage B (util/geom_io.py): what
geometry_from_nerf.process_view writes (util/geom.py:27-79) and what
datasets/nerf_shape.py:133-190 reads back."""
import json
import os

import numpy as np
import pytest

from nerfactor_b200 import synth
from nerfactor_b200.util import geom_io


def _fake_view(h=12, w=16, L=8, seed=0):
    rng = np.random.default_rng(seed)
    alpha = (rng.uniform(size=(h, w)) > 0.3).astype(np.float32) * rng.uniform(0.5, 1., size=(h, w)).astype(np.float32)
    xyz = rng.uniform(-1, 1, size=(h, w, 3)).astype(np.float32) * alpha[:, :, None]
    n = rng.standard_normal((h, w, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    lvis = rng.uniform(size=(h, w, L)).astype(np.float32) * alpha[:, :, None]
    return {'alpha': alpha, 'xyz': xyz, 'normal': n, 'lvis': lvis}


def _metadata(tmp_path, h, w):
    c2w = synth.look_at_c2w()
    md = {'imh': h, 'imw': w, 'cam_angle_x': synth.CAM_ANGLE_X,
          'cam_transform_mat': ','.join('%.17g' % x for x in np.asarray(c2w).reshape(-1))}
    d = tmp_path / 'test_000'
    d.mkdir()
    (d / 'metadata.json').write_text(json.dumps(md))
    return str(d / 'metadata.json')


def test_write_then_load_round_trip(tmp_path):
    h, w, L = 12, 16, 8
    buf = _fake_view(h, w, L)
    out_dir = str(tmp_path / 'geom' / 'test_000')
    assert not geom_io.view_done(out_dir)
    geom_io.write_view_buffers(buf, out_dir)
    assert geom_io.view_done(out_dir)
    from PIL import Image
    a8 = np.array(Image.open(os.path.join(out_dir, 'alpha.png')))
    assert a8.dtype == np.uint8 and np.array_equal(a8, (buf['alpha'].astype(np.float64) * 255).astype(np.uint8))
    meta = _metadata(tmp_path, h, w)
    id_, rayo, rayd, rgb, alpha, xyz, normal, lvis = geom_io.load_view(meta, out_dir, imh=h)
    assert id_ == 'test_000'
    assert rayo.shape == rayd.shape == (h, w, 3) and rayo.dtype == np.float32
    assert np.array_equal(xyz, buf['xyz']) and np.array_equal(lvis, buf['lvis'])
    assert np.allclose(normal, buf['normal'], atol=1e-6)
    assert np.abs(alpha - buf['alpha']).max() <= 1 / 255 and not rgb.any()
    # rays: same generator the device kernel is checked against (datasets/nerf.py:172-193)
    ro, rd = geom_io.gen_rays_host(np.asarray(synth.look_at_c2w()), synth.CAM_ANGLE_X, h, w)
    assert np.array_equal(rayo, ro.astype(np.float32)) and np.array_equal(rayd, rd.astype(np.float32))
    assert np.allclose(rayd[0, 0], (np.asarray(synth.look_at_c2w())[:3, :3] @ np.array(
        [(0 - .5 * w) / (.5 * w / np.tan(.5 * synth.CAM_ANGLE_X)),
         -(0 - .5 * h) / (.5 * w / np.tan(.5 * synth.CAM_ANGLE_X)), -1.])), atol=1e-6)


def test_load_resizes_and_train_mode(tmp_path):
    h, w, L = 16, 16, 4
    buf = _fake_view(h, w, L, seed=2)
    buf['alpha'][:] = 1.0          # resized normals of a masked view can cancel to zero length
    buf['xyz'] = np.abs(buf['xyz']) + 0.1
    out_dir = str(tmp_path / 'geom' / 'train_000')
    geom_io.write_view_buffers(buf, out_dir)
    meta = _metadata(tmp_path, h, w)
    rgba = (np.random.default_rng(3).uniform(size=(h, w, 4)) * 255).astype(np.uint8)
    from PIL import Image
    rgba_path = str(tmp_path / 'rgba.png')
    Image.fromarray(rgba).save(rgba_path)
    out = geom_io.load_view(meta, out_dir, imh=8, mode='train', rgba_path=rgba_path)
    _, rayo, rayd, rgb, alpha, xyz, normal, lvis = out
    assert rayo.shape == (8, 8, 3) and xyz.shape == (8, 8, 3) and lvis.shape == (8, 8, L)
    assert rgb.shape == (8, 8, 3) and alpha.shape == (8, 8)
    assert np.allclose(np.linalg.norm(normal, axis=2), 1., atol=1e-5)
    # INTER_AREA at exactly 2x = 2x2 box mean
    box = buf['lvis'].reshape(8, 2, 8, 2, L).mean(axis=(1, 3))
    assert np.allclose(lvis, np.clip(box, 0, 1), atol=1e-6)
    assert np.allclose(alpha, (rgba[:, :, 3] / 255.).reshape(8, 2, 8, 2).mean(axis=(1, 3)), atol=1e-6)
    with pytest.raises(ValueError):
        geom_io._write_png(np.full((2, 2), 1.5), str(tmp_path / 'x.png'))

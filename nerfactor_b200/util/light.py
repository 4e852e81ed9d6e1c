"""Mirror of nerfactor/util/light.py:27-66 (env-map visualisation) and the probe loading of
nerfactor/models/nerfactor.py:85-92, 169-179 (`_load_light`: read the Radiance .hdr, resample to
`light_h x 2 light_h` with TensorFlow's antialiased bilinear filter).  The loaded probes are what
`Model.call(relight_probes=True)` integrates against, so this is the data format on the input
side of the relighting path.
"""
import os

import numpy as np

from . import img as imgutil
from .io import sortglob


# ------------------------------------------------------------------ Radiance RGBE (.hdr)
def read_hdr(path):
    """xiuminglib io.hdr.read (cv2.imdecode + BGR->RGB): float32 [H,W,3], RGB order."""
    import cv2
    with open(path, 'rb') as h:
        buf = np.frombuffer(h.read(), np.uint8)
    bgr = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
    if bgr is None:
        raise IOError("cannot decode %s" % path)
    return np.ascontiguousarray(bgr[:, :, ::-1])


def write_hdr(rgb, outpath):
    """xiuminglib io.hdr.write."""
    import cv2
    rgb = np.asarray(rgb)
    assert rgb.dtype == np.float32, "Input must be float32"
    os.makedirs(os.path.dirname(os.path.abspath(outpath)), exist_ok=True)
    if not cv2.imwrite(outpath, np.ascontiguousarray(rgb[:, :, ::-1])):
        raise IOError("Writing HDR failed")


def load_light(path, light_h):
    """Model._load_light, nerfactor.py:169-179 -> float32 [light_h, 2 light_h (aspect kept), 3]."""
    ext = os.path.basename(path).split('.')[-1]
    if ext == 'hdr':
        arr = read_hdr(path)
    elif ext == 'exr':
        raise NotImplementedError("OpenEXR probes need the OpenEXR library (not in this image); "
                                  "convert to .hdr")
    else:
        raise NotImplementedError(ext)
    return imgutil.resize(arr.astype(np.float32), new_h=light_h)


def load_probes(test_envmap_dir, light_h):
    """nerfactor.py:85-92: every .hdr / .exr of `test_envmap_dir`, sorted, name -> [h, 2h, 3]."""
    from collections import OrderedDict
    probes = OrderedDict()
    if not test_envmap_dir or not os.path.isdir(test_envmap_dir):
        return probes
    for path in sortglob(test_envmap_dir, ext=('hdr', 'exr')):
        name = os.path.basename(path)[:-len('.hdr')]
        probes[name] = load_light(path, light_h)
    return probes


# ------------------------------------------------------------------------ visualisation
def one_hot_img(h, w, c, i, j):
    """util/tensor.py:57-64."""
    img = np.zeros((h, w, c), np.float32)
    img[i, j, :] = 1
    return img


def vis_light(light_probe, outpath=None, h=None):
    """util/light.py:27-48: resize (TF bilinear) -> gamma-4 tonemap -> uint8 (+ optional PNG)."""
    if hasattr(light_probe, 'detach'):
        light_probe = light_probe.detach().cpu().numpy()
    light_probe = np.asarray(light_probe, np.float32)
    if h is not None:
        light_probe = imgutil.resize(light_probe, new_h=h)
    img = imgutil.tonemap(light_probe, method='gamma', gamma=4)
    img_uint = imgutil.denormalize_float(img)
    if outpath is not None:
        imgutil.write_uint(img_uint, outpath)
    return img_uint


def vis_hdr_lights(lights_dir, vis_h=64):
    """util/light.py:51-58."""
    vis = {}
    for path in sortglob(lights_dir, ext='hdr'):
        vis[os.path.basename(path)[:-len('.hdr')]] = vis_light(read_hdr(path), h=vis_h)
    return vis


def vis_olat_lights(orig_h=16, vis_h=64):
    """util/light.py:61-66."""
    vis = {}
    for i in range(orig_h):
        for j in range(2 * orig_h):
            vis['%04d-%04d' % (i, j)] = vis_light(one_hot_img(orig_h, 2 * orig_h, 3, i, j), h=vis_h)
    return vis

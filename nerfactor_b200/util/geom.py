"""Mirror of nerfactor/util/geom.py: the geometry-buffer writers (geom.py:27-79; implemented in
util/geom_io.py) and the direction algebra of the learned BRDF (geom.py:96-192), the latter as
torch functions that keep the reference's custom gradients (autodiff.py).  Inside the models
`gen_world2local` / `dir2rusink` are fused into the BRDF kernel prologue (csrc/nf_mlp_tc.cu)."""
import math

import torch

from ..autodiff import gen_world2local, dir2rusink  # noqa: F401
from .geom_io import write_alpha, write_xyz, write_normal  # noqa: F401
from . import geom_io as _geom_io


def write_lvis(lvis, fps, out_dir):
    """geom.py:27-44 (the per-light video is written by models/_visualize when cv2 is present;
    `fps` is accepted for signature compatibility)."""
    return _geom_io.write_lvis(lvis, out_dir)


def rad2deg(rad):
    return 180. / math.pi * rad


def slerp(p0, p1, t):
    """Spherical interpolation of two vectors (geom.py:100-116)."""
    p0, p1 = torch.as_tensor(p0, dtype=torch.float32), torch.as_tensor(p1, dtype=torch.float32)
    cos_omega = torch.sum(p0 / torch.linalg.norm(p0) * (p1 / torch.linalg.norm(p1)))
    omega = torch.acos(torch.clamp(cos_omega, -1., 1.))
    so = torch.sin(omega)
    if float(so) == 0.:
        return (1. - t) * p0 + t * p1
    return torch.sin((1. - t) * omega) / so * p0 + torch.sin(t * omega) / so * p1

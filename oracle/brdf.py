"""Oracle restatement of brdf/ and nerfactor/util/geom.py (test infrastructure)."""
import math

import numpy as np
import torch

from . import tfops
from .tfops import safe_l2_normalize, divide_no_nan


# ----------------------------------------------------------------- light grid

def sph2cart(pts_r_lat_lng):
    """third_party/xiuminglib/xiuminglib/geometry/sph.py:184-193 ('lat-lng')."""
    r, lat, lng = pts_r_lat_lng[:, 0], pts_r_lat_lng[:, 1], pts_r_lat_lng[:, 2]
    z = r * np.sin(lat)
    x = r * np.cos(lat) * np.cos(lng)
    y = r * np.cos(lat) * np.sin(lng)
    return np.stack((x, y, z), axis=-1)


def gen_light_xyz(envmap_h, envmap_w, envmap_radius=1e2):
    """brdf/renderer.py:184-219. fp64 NumPy; callers cast to fp32
    (nerfactor/models/shape.py:75-76)."""
    lat_step_size = np.pi / (envmap_h + 2)
    lng_step_size = 2 * np.pi / (envmap_w + 2)
    lats = np.linspace(
        np.pi / 2 - lat_step_size, -np.pi / 2 + lat_step_size, envmap_h)
    lngs = np.linspace(
        np.pi - lng_step_size, -np.pi + lng_step_size, envmap_w)
    lngs, lats = np.meshgrid(lngs, lats)
    rlatlngs = np.dstack((envmap_radius * np.ones_like(lats), lats, lngs))
    rlatlngs = rlatlngs.reshape(-1, 3)
    xyz = sph2cart(rlatlngs).reshape(envmap_h, envmap_w, 3)
    sin_colat = np.sin(np.pi / 2 - lats)
    areas = 4 * np.pi * sin_colat / np.sum(sin_colat)
    return xyz, areas


# ------------------------------------------------------------ microfacet BRDF

class Microfacet:
    """brdf/microfacet/microfacet.py:21-111 (GGX D, Schlick F, view-side G1)."""

    def __init__(self, default_rough=0.3, lambert_only=False, f0=0.91):
        self.default_rough = default_rough
        self.lambert_only = lambert_only
        self.f0 = f0

    def __call__(self, pts2l, pts2c, normal, albedo=None, rough=None):
        dt = pts2l.dtype
        n_pts = pts2c.shape[0]
        if albedo is None:
            albedo = torch.ones((n_pts, 3), dtype=dt)
        if rough is None:
            rough = self.default_rough * torch.ones((n_pts, 1), dtype=dt)
        # microfacet.py:46-49
        pts2l = safe_l2_normalize(pts2l, 2)
        pts2c = safe_l2_normalize(pts2c, 1)
        normal = safe_l2_normalize(normal, 1)
        # microfacet.py:51-61
        h = pts2l + pts2c[:, None, :]
        h = safe_l2_normalize(h, 2)
        f = self._get_f(pts2l, h)
        alpha = rough ** 2
        d = self._get_d(h, normal, alpha=alpha)
        g = self._get_g(pts2c, h, normal, alpha=alpha)
        l_dot_n = torch.einsum('ijk,ik->ij', pts2l, normal)
        v_dot_n = torch.einsum('ij,ij->i', pts2c, normal)
        denom = 4 * torch.abs(l_dot_n) * torch.abs(v_dot_n)[:, None]
        microfacet = divide_no_nan(f * g * d, denom)
        brdf_glossy = microfacet[:, :, None].repeat(1, 1, 3)
        # microfacet.py:63-72
        lambert = albedo / math.pi
        brdf_diffuse = lambert[:, None, :].expand_as(brdf_glossy)
        if self.lambert_only:
            return brdf_diffuse
        return brdf_glossy + brdf_diffuse

    @staticmethod
    def _get_g(v, m, n, alpha):
        """microfacet.py:74-90."""
        cos_theta_v = torch.einsum('ij,ij->i', n, v)
        cos_theta = torch.einsum('ijk,ik->ij', m, v)
        denom = cos_theta_v[:, None]
        div = divide_no_nan(cos_theta, denom)
        chi = torch.where(div > 0, torch.ones_like(div), torch.zeros_like(div))
        cos_theta_v_sq = torch.clamp(cos_theta_v ** 2, 0., 1.)
        tan_theta_v_sq = divide_no_nan(1 - cos_theta_v_sq, cos_theta_v_sq)
        tan_theta_v_sq = torch.clamp(tan_theta_v_sq, min=0.)
        denom = 1 + torch.sqrt(1 + alpha ** 2 * tan_theta_v_sq[:, None])
        return divide_no_nan(chi * 2, denom)

    @staticmethod
    def _get_d(m, n, alpha):
        """microfacet.py:92-104 (note alpha**2 again on top of alpha=rough**2)."""
        cos_theta_m = torch.einsum('ijk,ik->ij', m, n)
        chi = torch.where(
            cos_theta_m > 0, torch.ones_like(cos_theta_m),
            torch.zeros_like(cos_theta_m))
        cos_theta_m_sq = cos_theta_m ** 2
        tan_theta_m_sq = divide_no_nan(1 - cos_theta_m_sq, cos_theta_m_sq)
        denom = math.pi * cos_theta_m_sq ** 2 * (alpha ** 2 + tan_theta_m_sq) ** 2
        return divide_no_nan(alpha ** 2 * chi, denom)

    def _get_f(self, l, m):
        """microfacet.py:106-111."""
        cos_theta = torch.einsum('ijk,ijk->ij', l, m)
        return self.f0 + (1 - self.f0) * (1 - cos_theta) ** 5


# ------------------------------------------------- local frames & Rusinkiewicz

def gen_world2local(normal, eps=1e-6):
    """nerfactor/util/geom.py:119-149: rows (t, b, n)."""
    normal = safe_l2_normalize(normal, 1)
    z = torch.tensor((0., 0., 1.), dtype=normal.dtype) + eps  # all three comps
    z = z[None, :].expand(normal.shape[0], 3)
    t = torch.linalg.cross(normal, z)
    t = safe_l2_normalize(t, 1)
    b = torch.linalg.cross(normal, t)
    b = safe_l2_normalize(b, 1)
    return torch.stack((t, b, normal), dim=1)


class _SafeAcos(torch.autograd.Function):
    """nerfactor/util/math.py:42-60: acos(clip(x)) with the reference's custom gradient
    -1 / (sqrt(1 - x_clip^2 + eps) + eps), eps = 1e-6."""

    @staticmethod
    def forward(ctx, x):
        xc = torch.clamp(x, -1., 1.)
        ctx.save_for_backward(xc)
        return torch.acos(xc)

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        return dy * (-1. / (torch.sqrt(1. - xc ** 2 + 1e-6) + 1e-6))


class _SafeAtan2(torch.autograd.Function):
    """nerfactor/util/math.py:24-39: atan2(x, y) with denominators x^2 + y^2 + eps."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return torch.atan2(x, y)

    @staticmethod
    def backward(ctx, dz):
        x, y = ctx.saved_tensors
        denom = x ** 2 + y ** 2 + 1e-6
        return dz * (y / denom), dz * (-x / denom)


def safe_acos(x):
    return _SafeAcos.apply(x)


def safe_atan2(x, y):
    return _SafeAtan2.apply(x, y)


def dir2rusink(a, b):
    """nerfactor/util/geom.py:152-192. Returns (phi_d, theta_h, theta_d)."""
    dt = a.dtype
    a = safe_l2_normalize(a, 1)
    b = safe_l2_normalize(b, 1)
    h = safe_l2_normalize((a + b) / 2, 1)
    theta_h = safe_acos(h[:, 2])
    phi_h = safe_atan2(h[:, 1], h[:, 0])
    binormal = torch.tensor((0., 1., 0.), dtype=dt)
    normal = torch.tensor((0., 0., 1.), dtype=dt)

    def rot_vec(vector, axis, angle):
        cos_ang = torch.cos(angle).reshape(-1)
        sin_ang = torch.sin(angle).reshape(-1)
        vector = vector.reshape(-1, 3)
        axis = axis.reshape(-1, 3)
        return vector * cos_ang[:, None] + \
            axis * (vector @ axis.t()) * (1 - cos_ang)[:, None] + \
            torch.linalg.cross(
                axis.expand(vector.shape[0], 3), vector) * sin_ang[:, None]

    diff = rot_vec(rot_vec(b, normal, -phi_h), binormal, -theta_h)
    diff0, diff1, diff2 = diff[:, 0], diff[:, 1], diff[:, 2]
    theta_d = safe_acos(diff2)
    phi_d = tfops.floormod(safe_atan2(diff1, diff0), math.pi)
    return torch.stack((phi_d, theta_h, theta_d), dim=1)

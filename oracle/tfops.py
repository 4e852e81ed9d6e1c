"""TensorFlow-2.2 op semantics restated on PyTorch-CPU (test infrastructure).

Each helper mirrors one `tf.*` call used on the hot path.
"""
import math

import torch


def l2_normalize(x, axis, epsilon=1e-12):
    """tf.linalg.l2_normalize: x * rsqrt(max(sum(x^2), eps)); eps compared with
    the SQUARED norm (used at nerfactor/util/math.py:63-64 with eps=1e-6 and at
    nerfactor/geometry_from_nerf.py:100,198,297 with the default)."""
    sq = torch.sum(x * x, dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))


def safe_l2_normalize(x, axis, eps=1e-6):
    """nerfactor/util/math.py:63-64."""
    return l2_normalize(x, axis, epsilon=eps)


def divide_no_nan(a, b):
    """tf.math.divide_no_nan: 0 where b == 0 (brdf/microfacet/microfacet.py:61)."""
    a, b = torch.broadcast_tensors(a, b)
    safe = torch.where(b == 0, torch.ones_like(b), b)
    return torch.where(b == 0, torch.zeros_like(a), a / safe)


def cumprod_exclusive(x):
    """tf.math.cumprod(x, axis=-1, exclusive=True): out[i] = prod_{j<i} x[j]."""
    cp = torch.cumprod(x, dim=-1)
    return torch.cat((torch.ones_like(cp[..., :1]), cp[..., :-1]), dim=-1)


def safe_cumprod(x, eps=1e-6):
    """nerfactor/util/math.py:67-68."""
    return cumprod_exclusive(x + eps)


def searchsorted_right(cdf, u):
    """tf.searchsorted(cdf, u, side='right'): first i with cdf[i] > u."""
    return torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True)


def floormod(x, y):
    """tf.math.floormod: result takes the sign of the divisor."""
    return x - torch.floor(x / y) * y


def linspace(a, b, n, dtype):
    """tf.linspace(a, b, n) in the given dtype: a + i*(b-a)/(n-1)."""
    if n == 1:
        return torch.tensor([a], dtype=dtype)
    i = torch.arange(n, dtype=dtype)
    step = (torch.tensor(b, dtype=dtype) - torch.tensor(a, dtype=dtype)) / (n - 1)
    return torch.tensor(a, dtype=dtype) + i * step


PI = math.pi

"""tf.math.* used on the path (see the package docstring: test infrastructure)."""
import torch

sin, cos, exp, log, sqrt, abs, square, acos = (  # noqa: A001
    torch.sin, torch.cos, torch.exp, torch.log, torch.sqrt, torch.abs, torch.square, torch.acos)
sigmoid = torch.sigmoid
rsqrt = torch.rsqrt


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, tuple) and hasattr(x, 'numpy'):
        x = list(x)
    if isinstance(x, float):
        return torch.as_tensor(x, dtype=torch.float32)
    if isinstance(x, int):
        return torch.as_tensor(x, dtype=torch.int32)
    return torch.as_tensor(x)


def atan2(y, x):
    return torch.atan2(_t(y), _t(x))


def pow(x, y):  # noqa: A001
    return torch.pow(_t(x), y)


def minimum(a, b):
    a, b = _t(a), _t(b)
    return torch.minimum(a, b.to(a.dtype))


def maximum(a, b):
    a, b = _t(a), _t(b)
    return torch.maximum(a, b.to(a.dtype))


def divide_no_nan(x, y):
    """x / y, 0 where y == 0."""
    x, y = _t(x), _t(y)
    safe = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(x * y), x / safe)


def floormod(x, y):
    """Result has the sign of the divisor (Python %)."""
    return torch.remainder(_t(x), y)


def cumprod(x, axis=0, exclusive=False, reverse=False):
    assert not reverse
    out = torch.cumprod(x, dim=axis)
    if exclusive:
        ones = torch.ones_like(x.narrow(axis, 0, 1))
        out = torch.cat((ones, out.narrow(axis, 0, x.shape[axis] - 1)), dim=axis)
    return out


def l2_normalize(x, axis=None, epsilon=1e-12):
    sq = torch.sum(x * x, dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))

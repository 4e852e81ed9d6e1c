"""tf.linalg.* used on the path."""
import torch

from .math import l2_normalize  # noqa: F401


def norm(tensor, ord='euclidean', axis=None, keepdims=False):  # noqa: A002
    assert ord in ('euclidean', 2)
    return torch.sqrt(torch.sum(tensor * tensor, dim=axis, keepdim=keepdims)) if axis is not None \
        else torch.sqrt(torch.sum(tensor * tensor))


def cross(a, b):
    return torch.cross(a, b, dim=-1)

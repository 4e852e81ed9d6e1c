"""Mirror of nerfactor/util/math.py:63-64 for small [N,3] host-side glue."""
import torch


def safe_l2_normalize(x, axis=None, eps=1e-6):
    sq = torch.sum(x * x, dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=eps))

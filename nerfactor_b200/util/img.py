"""Mirror of nerfactor/util/img.py:76-95, 140-163 (loss-side glue on [N,3])."""
import torch


def alpha_blend(tensor1, alpha, tensor2=None):
    if tensor2 is None:
        tensor2 = torch.zeros_like(tensor1)
    return tensor1 * alpha + tensor2 * (1. - alpha)


def linear2srgb(tensor_0to1):
    x = torch.clamp(tensor_0to1, 0., 1.)
    return torch.where(x <= 0.0031308, x * 12.92, 1.055 * torch.pow(x, 1 / 2.4) - 0.055)

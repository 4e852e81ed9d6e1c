"""tf.image.resize(method='bilinear', antialias=True): ScaleAndTranslate spans (half-pixel centres,
triangle kernel stretched by max(in / out, 1), renormalised), separable, float32."""
import numpy as np
import torch


def _weights(in_size, out_size, antialias):
    inv_scale = in_size / out_size
    ks = max(inv_scale, 1.) if antialias else 1.
    mat = np.zeros((out_size, in_size), np.float32)
    for x in range(out_size):
        s = (x + 0.5) * inv_scale
        lo = min(max(int(np.ceil(s - ks - 0.5)), 0), in_size - 1)
        hi = min(max(int(np.floor(s + ks - 0.5)), 0), in_size - 1)
        src = np.arange(lo, hi + 1)
        w = np.maximum(0., 1. - np.abs((src + 0.5 - s) / ks))
        tot = w.sum()
        mat[x, src] = w / tot if abs(tot) >= 1000. * np.finfo(np.float32).tiny else w
    return torch.from_numpy(mat)


def resize(images, size, method='bilinear', antialias=False, **_):
    assert method == 'bilinear'
    x = torch.as_tensor(images).to(torch.float32)
    nh, nw = int(size[0]), int(size[1])
    squeeze = x.dim() == 3
    if squeeze:
        x = x[None]
    wy, wx = _weights(x.shape[1], nh, antialias), _weights(x.shape[2], nw, antialias)
    x = torch.einsum('oh,nhwc->nowc', wy, x)
    x = torch.einsum('pw,nowc->nopc', wx, x)
    return x[0] if squeeze else x

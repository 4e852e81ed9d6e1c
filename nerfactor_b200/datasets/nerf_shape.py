"""Mirror of nerfactor/datasets/nerf_shape.py:72-121: turns one view's buffers
(`rayo, rayd, rgb [H,W,3]`, `alpha [H,W]`, `xyz, normal [H,W,3]`, `lvis [H,W,L]`) into the
9-tuple batch `Model.call` consumes.  Works on torch tensors on any device."""
import torch


def sample_rays(rayo, rayd, rgb, alpha, xyz, normal, lvis, mode='train', bs=1024,
                alpha_thres=0.9, always_all_rays=False, generator=None):
    """nerf_shape.py:84-121.  vali / test (or always_all_rays): every ray, row-major
    (ray n = y * W + x).  train: `bs` uniform draws WITH replacement from the pixels whose
    alpha exceeds `alpha_thres` (all pixels if alpha_thres is None)."""
    h, w = rgb.shape[0], rgb.shape[1]
    if mode in ('vali', 'test') or always_all_rays:
        f = lambda t, c: t.reshape(-1, c)
        return (f(rayo, 3), f(rayd, 3), f(rgb, 3), f(alpha, 1), f(xyz, 3), f(normal, 3),
                lvis.reshape(h * w, -1))
    if alpha_thres is None:
        coords = torch.arange(h * w, device=rgb.device)
    else:
        coords = torch.nonzero(alpha.reshape(-1) > alpha_thres, as_tuple=False)[:, 0]
    if coords.numel() == 0:
        raise ValueError("no foreground pixel above alpha_thres in this view")
    sel = coords[torch.randint(0, coords.numel(), (bs,), device=coords.device,
                               generator=generator)]
    g = lambda t, c: t.reshape(h * w, c).index_select(0, sel)
    return (g(rayo, 3), g(rayd, 3), g(rgb, 3), g(alpha, 1), g(xyz, 3), g(normal, 3),
            lvis.reshape(h * w, -1).index_select(0, sel))


def make_batch(id_, hw, rays):
    """nerf_shape.py:72-82: (id_, hw, rayo, rayd, rgb, alpha, xyz, normal, lvis); id_ / hw are
    per-view scalars here (the reference tiles them per ray only for tf.distribute)."""
    return (id_, hw) + tuple(rays)

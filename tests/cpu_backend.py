"""TEST DOUBLE for `nerfactor_b200._lib` (the ctypes layer over libnerfactor_b200.so).

The product has no CPU path: without a B200 `_lib.Context()` raises.  To exercise the HOST logic
above the C ABI in the CPU test-suite (dataset -> Model.call glue -> loss -> Trainer bookkeeping
-> checkpoints -> vis_batch -> the three drop-in scripts), `install(monkeypatch)` swaps every
`_lib` entry point the host code calls for a small PyTorch-CPU implementation built on the oracle
(`oracle/`, itself test infrastructure).  Nothing outside `tests/` imports this module, and the
GPU tests (`-m gpu`) never install it: they run the real kernels and compare them with the oracle.
"""
import numpy as np
import torch

from oracle import brdf as obrdf, networks as onets, stage_a, stage_b, tfops
from nerfactor_b200 import _lib


class FakeContext:
    def __init__(self, device=None):
        self.device = torch.device('cpu')
        self.launches = 0
        self.sm_count = 148


class FakePackedMlp:
    def __init__(self, ctx, kind, layers, skip_at, out_act, n_freqs_a=0, n_freqs_b=0, z_dim=0,
                 rgb=None):
        self.ctx, self.kind = ctx, kind
        layers = [(np.asarray(w, np.float32), np.asarray(b, np.float32)) for w, b in layers]
        depth = len(layers) - 1
        self.trunk = {'layers': layers[:-1], 'act': ['relu'] * depth, 'skip_at': [skip_at]}
        self.head = {'layers': layers[-1:], 'act': [out_act]}
        self.n_freqs_a, self.n_freqs_b, self.z_dim = n_freqs_a, n_freqs_b, z_dim
        self.out_dim, self.depth, self.width = layers[-1][0].shape[1], depth, layers[0][0].shape[1]
        self.has_rgb = rgb is not None

    def __call__(self, x):
        return onets.mlp_forward(self.head, onets.mlp_forward(self.trunk, x))


_ctx = None


def default_context():
    global _ctx
    if _ctx is None:
        _ctx = FakeContext()
    return _ctx


def point_mlp_fwd(ctx, mlp, xyz, xyz_scale=1.0, precision='fp32'):
    ctx.launches += 1
    return mlp(onets.embed(xyz * xyz_scale, mlp.n_freqs_a))


def lvis_fwd(ctx, mlp, xyz, lxyz, xyz_scale=1.0, precision='f16'):
    ctx.launches += 1
    lxyz = lxyz.reshape(-1, 3)
    n, L = xyz.shape[0], lxyz.shape[0]
    if n == 0:
        return torch.zeros((0, L))
    surf2l = tfops.safe_l2_normalize(lxyz[None] - xyz[:, None], 2)         # shape.py:128-135
    e_xyz = onets.embed(xyz * xyz_scale, mlp.n_freqs_a)[:, None, :].expand(n, L, -1)
    x = torch.cat((e_xyz, onets.embed(surf2l, mlp.n_freqs_b)), -1).reshape(n * L, -1)
    return mlp(x).reshape(n, L)


def brdf_learned_fwd(ctx, mlp, xyz, normal, cam, z, lxyz, precision='f16'):
    raise NotImplementedError("learned-BRDF lobe: covered by the GPU parity tests only")


def _pair_terms(xyz, normal, cam, albedo, lvis, lxyz, lareas, rough, spec, f0, spec_scale):
    lxyz, lareas = lxyz.reshape(-1, 3), lareas.reshape(-1)
    l = tfops.safe_l2_normalize(lxyz[None] - xyz[:, None], 2)
    v = tfops.safe_l2_normalize(cam - xyz, 1)
    n = tfops.safe_l2_normalize(normal, 1)
    if spec is None:
        brdf = obrdf.Microfacet(f0=f0)(l, v, n, albedo=albedo, rough=rough)
    else:
        brdf = albedo[:, None, :] / np.pi + (spec * spec_scale)[:, :, None]
    cos = torch.einsum('ijk,ik->ij', l, n)
    w = (cos > 0).to(cos.dtype) * lvis * cos * lareas[None, :]              # nerfactor.py:325-335
    return brdf * w[:, :, None]                                            # [n, L, 3]


def _tonemap(rgb, linear2srgb):
    rgb = torch.clamp(rgb, 0., 1.)
    return stage_b.linear2srgb(rgb) if linear2srgb else rgb


def integrate_fwd(ctx, xyz, normal, cam, albedo, lvis, lxyz, lareas, light, rough=None,
                  spec=None, light_idx=None, f0=0.04, spec_scale=1.0, linear2srgb=True):
    ctx.launches += 1
    c = _pair_terms(xyz, normal, cam, albedo, lvis, lxyz, lareas, rough, spec, f0, spec_scale)
    if light_idx is not None:
        light = light[:, light_idx.long(), :]
    return _tonemap(torch.einsum('nlc,elc->nec', c, light), linear2srgb)


def integrate_olat_fwd(ctx, xyz, normal, cam, albedo, lvis, lxyz, lareas, olat_inten, ambient,
                       rough=None, spec=None, f0=0.04, spec_scale=1.0, linear2srgb=True):
    ctx.launches += 1
    c = _pair_terms(xyz, normal, cam, albedo, lvis, lxyz, lareas, rough, spec, f0, spec_scale)
    return _tonemap(ambient * c.sum(1, keepdim=True) + olat_inten * c, linear2srgb)


def gen_rays(ctx, c2w, cam_angle_x, h, w, normalize=False):
    ctx.launches += 1
    rayo, rayd = stage_a.gen_rays(np.asarray(c2w, np.float64).reshape(4, 4), cam_angle_x, h, w)
    rayo = torch.as_tensor(rayo.reshape(-1, 3).astype(np.float32))
    rayd = torch.as_tensor(rayd.reshape(-1, 3).astype(np.float32))
    return rayo, (tfops.l2_normalize(rayd, 1) if normalize else rayd)


def gen_z(ctx, near, far, n_samples, n_rays, lin_in_disp=False, perturb_u=None):
    ctx.launches += 1
    return stage_a.gen_z(near, far, n_samples, n_rays, lin_in_disp, perturb_u)


def _sigma_raw(mlp, rayo, rayd, z):
    pts = (rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]).reshape(-1, 3)
    return pts, lambda p: torch.relu(mlp(onets.embed(p, mlp.n_freqs_a)))


def sigma_fwd(ctx, mlp, rayo, rayd, z, bbox=None, precision='f16'):
    ctx.launches += 1
    pts, f = _sigma_raw(mlp, rayo, rayd, z)
    sigma = f(pts)[:, 0] * stage_a.check_bounds(pts, bbox).to(pts.dtype)
    return sigma.reshape(z.shape)


def sigma_normal_fwd(ctx, mlp, rayo, rayd, z, bbox=None, precision='fp32'):
    ctx.launches += 1
    pts, f = _sigma_raw(mlp, rayo, rayd, z)
    with torch.enable_grad():
        p = pts.detach().clone().requires_grad_(True)
        s = f(p)
        (g,) = torch.autograd.grad(s.sum(), p)
    inb = stage_a.check_bounds(pts, bbox).to(pts.dtype)
    normal = -tfops.l2_normalize(g, 1) * inb[:, None]                      # gfn.py:289-305
    return (s.detach()[:, 0] * inb).reshape(z.shape), normal.reshape(z.shape + (3,))


def nerf_fwd(ctx, mlp, rayo, rayd, z, precision='f16'):
    raise NotImplementedError("NeRF colour branch: covered by the GPU parity tests only")


def composite(ctx, sigma, z, rayo, rayd, normal=None, want_weights=True, want_surf=True):
    ctx.launches += 1
    w = stage_a.accumulate_sigma(sigma, z, rayd)
    occu, depth = w.sum(-1), (w * z).sum(-1)
    surf = rayo + rayd * depth[:, None] if want_surf else None
    en = (w[:, :, None] * normal).sum(1) if normal is not None else None
    return (w if want_weights else None), occu, depth, surf, en


def gen_z_fine(ctx, z_coarse, weights, n_fine):
    ctx.launches += 1
    return stage_a.gen_z_fine(z_coarse, weights, n_fine)


def lvis_rays(ctx, surf, normal, lxyz):
    """geometry_from_nerf.py:196-215."""
    ctx.launches += 1
    n, L = surf.shape[0], lxyz.shape[0]
    d = tfops.l2_normalize(lxyz[None] - surf[:, None], 2)
    fl = (torch.einsum('ijk,ik->ij', d, normal) > 0).to(torch.uint8)
    return surf[:, None, :].expand(n, L, 3).reshape(-1, 3).contiguous(), d.reshape(-1, 3), fl


def _act(name, y):
    return onets._act(name, y)


def dense_fwd(ctx, x1, x2, w, b, act, precision='fp32'):
    ctx.launches += 1
    x = x1 if x2 is None else torch.cat((x1, x2), 1)
    return _act(act, x @ w + b)


def dense_bwd(ctx, x1, x2, w, y, dy, act, need_dx1, need_dx2, precision='fp32'):
    ctx.launches += 1
    x = x1 if x2 is None else torch.cat((x1, x2), 1)
    if act == 'relu':
        dz = dy * (y > 0).to(dy.dtype)
    elif act == 'sigmoid':
        dz = dy * y * (1 - y)
    elif act == 'softplus':
        dz = dy * (1 - torch.exp(-y))
    else:
        dz = dy
    dx = dz @ w.t()
    k1 = x1.shape[1]
    return (dx[:, :k1].contiguous() if need_dx1 else None,
            dx[:, k1:].contiguous() if (need_dx2 and x2 is not None) else None,
            x.t() @ dz, dz.sum(0))


def _act_fwd(z, act):
    if act == 'relu':
        return torch.relu(z)
    if act == 'sigmoid':
        return torch.sigmoid(z)
    if act == 'softplus':
        return torch.nn.functional.softplus(z)
    return z


def mlp_chain_fwd(ctx, x, ws, bs, acts, skip_layer, precision='bf16'):
    """CPU double of nf_mlp_chain_fwd: the `workspace` is the list of layer inputs / outputs."""
    ctx.launches += 1
    h, saved = x, []
    for l, (w, b, a) in enumerate(zip(ws, bs, acts)):
        xin = torch.cat((h, x), 1) if (skip_layer and l == skip_layer) else h
        h = _act_fwd(xin @ w + b, a)
        saved.append((xin, h))
    return h, saved


def mlp_chain_bwd(ctx, ws, bs, acts, skip_layer, in_dim, y, dy, work, need_dx, precision='bf16'):
    ctx.launches += 1
    dws, dbs = [None] * len(ws), [None] * len(ws)
    dh, dx_skip = dy, None
    for l in range(len(ws) - 1, -1, -1):
        xin, out = work[l]
        a = acts[l]
        dz = dh * ((out > 0).to(dh.dtype) if a == 'relu' else out * (1 - out) if a == 'sigmoid'
                   else (1 - torch.exp(-out)) if a == 'softplus' else 1.)
        dws[l], dbs[l] = xin.t() @ dz, dz.sum(0)
        dxin = dz @ ws[l].t()
        if skip_layer and l == skip_layer:
            dh, dx_skip = dxin[:, :xin.shape[1] - in_dim], dxin[:, xin.shape[1] - in_dim:]
        else:
            dh = dxin
    dx = (dh + dx_skip if dx_skip is not None else dh).contiguous() if need_dx else None
    return dx, dws, dbs


def adam_amsgrad_step(ctx, param, grad, m, v, vhat, lr, step, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam(amsgrad=True), trainvali.py:110-127."""
    ctx.launches += 1
    m.mul_(beta1).add_(grad, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    torch.maximum(vhat, v, out=vhat)
    lr_t = lr * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    param.sub_(lr_t * m / (vhat.sqrt() + eps))


def microfacet_brdf_fwd(ctx, pts2l, pts2c, normal, albedo=None, rough=None, default_rough=0.3,
                        lambert_only=False, f0=0.91):
    ctx.launches += 1
    m = obrdf.Microfacet(default_rough=default_rough, lambert_only=lambert_only, f0=f0)
    return m(pts2l, pts2c, normal, albedo, None if rough is None else rough.reshape(-1, 1))


def raymarch_depth_normal_fwd(ctx, mlp_coarse, mlp_fine, rayo, rayd, near, far, n_coarse, n_fine,
                              lin_in_disp=False, bbox=None, precision='f16e'):
    n = rayo.shape[0]
    z = gen_z(ctx, near, far, n_coarse, n, lin_in_disp, None)
    sigma = sigma_fwd(ctx, mlp_coarse, rayo, rayd, z, bbox, precision)
    w, _, _, _, _ = composite(ctx, sigma, z, rayo, rayd, want_surf=False)
    z = gen_z_fine(ctx, z, w, n_fine)
    sigma, normal = sigma_normal_fwd(ctx, mlp_fine, rayo, rayd, z, bbox, 'fp32')
    _, occu, depth, _, en = composite(ctx, sigma, z, rayo, rayd, normal=normal,
                                      want_weights=False, want_surf=False)
    return occu, depth, en


def raymarch_lvis_fwd(ctx, mlp_coarse, mlp_fine, surf, normal, lxyz, lvis_near, lvis_far,
                      n_coarse, n_fine, lin_in_disp=False, bbox=None, precision='f16e'):
    m, L = surf.shape[0], lxyz.shape[0]
    rayo, rayd, fl = lvis_rays(ctx, surf, normal, lxyz)
    idx = torch.nonzero(fl.reshape(-1), as_tuple=False)[:, 0]
    lvis = torch.zeros((m * L,), dtype=torch.float32)
    if idx.numel():
        o, d = rayo[idx].contiguous(), rayd[idx].contiguous()
        z = gen_z(ctx, lvis_near, lvis_far, n_coarse, o.shape[0], lin_in_disp, None)
        sigma = sigma_fwd(ctx, mlp_coarse, o, d, z, bbox, precision)
        w, _, _, _, _ = composite(ctx, sigma, z, o, d, want_surf=False)
        z = gen_z_fine(ctx, z, w, n_fine)
        sigma = sigma_fwd(ctx, mlp_fine, o, d, z, bbox, precision)
        _, occu, _, _, _ = composite(ctx, sigma, z, o, d, want_weights=False, want_surf=False)
        lvis[idx] = 1. - occu
    return lvis.reshape(m, L)


def lvis_dirs_fwd(ctx, mlp, xyz, xyz_dir, lxyz, xyz_scale=1.0, precision='f16'):
    ctx.launches += 1
    lxyz = lxyz.reshape(-1, 3)
    n, L = xyz.shape[0], lxyz.shape[0]
    if n == 0:
        return torch.zeros((0, L))
    surf2l = tfops.safe_l2_normalize(lxyz[None, :, :] - xyz_dir[:, None, :], 2)
    e_xyz = onets.embed(xyz * xyz_scale, mlp.n_freqs_a)
    x = torch.cat((e_xyz[:, None, :].expand(n, L, e_xyz.shape[1]).reshape(n * L, -1),
                   onets.embed(surf2l.reshape(-1, 3), mlp.n_freqs_b)), -1)
    return mlp(x).reshape(n, L)


def stageB_fused_fwd(ctx, mlp_lvis, xyz, normal, cam, albedo, lxyz, lareas, light, rough=None,
                     z=None, mlp_brdf=None, light_idx=None, f0=0.04, spec_scale=1.0, xyz_scale=1.0,
                     linear2srgb=True, precision='f16', want_lvis=False, all_lights=False):
    lvis = lvis_fwd(ctx, mlp_lvis, xyz, lxyz, xyz_scale, precision)
    spec = None if z is None else brdf_learned_fwd(ctx, mlp_brdf, xyz, normal, cam, z, lxyz, precision)
    rgb = integrate_fwd(ctx, xyz, normal, cam, albedo, lvis, lxyz, lareas, light,
                        rough=None if rough is None else rough.reshape(-1, 1), spec=spec,
                        light_idx=light_idx, f0=f0, spec_scale=spec_scale, linear2srgb=linear2srgb)
    return rgb, (lvis if want_lvis else None)


_PATCHED = ('default_context', 'microfacet_brdf_fwd', 'stageB_fused_fwd', 'lvis_dirs_fwd', 'raymarch_depth_normal_fwd',
            'raymarch_lvis_fwd', 'point_mlp_fwd', 'lvis_fwd', 'brdf_learned_fwd', 'integrate_fwd',
            'integrate_olat_fwd', 'gen_rays', 'gen_z', 'sigma_fwd', 'sigma_normal_fwd',
            'nerf_fwd', 'composite', 'gen_z_fine', 'lvis_rays', 'dense_fwd', 'dense_bwd',
            'mlp_chain_fwd', 'mlp_chain_bwd', 'adam_amsgrad_step')


def install(monkeypatch):
    """Patches `nerfactor_b200._lib` for the duration of one test."""
    global _ctx
    _ctx = None
    g = globals()
    for name in _PATCHED:
        monkeypatch.setattr(_lib, name, g[name])
    monkeypatch.setattr(_lib, 'Context', FakeContext)
    monkeypatch.setattr(_lib, 'PackedMlp', FakePackedMlp)
    monkeypatch.setattr(_lib, '_default_ctx', None)
    return default_context()

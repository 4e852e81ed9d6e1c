"""Oracle restatement of Stage A: camera rays -> NeRF sigma-MLP march ->
surface / normal / light visibility (nerfactor/geometry_from_nerf.py,
nerfactor/models/nerf.py, nerfactor/datasets/nerf.py). Test infrastructure.

`nerf` parameter dicts: {'coarse_enc', 'coarse_sigma_out', 'fine_enc',
'fine_sigma_out'} (oracle.networks mlp dicts), n_freqs_xyz = 10.
"""
import numpy as np
import torch

from . import networks as nets
from . import brdf as brdfmod
from . import tfops
from .tfops import l2_normalize


# ------------------------------------------------------------------ ray gen

def gen_rays(to_world, angle_x, imh, imw, sps=1):
    """nerfactor/datasets/nerf.py:172-193 (ndc=False): fp64 NumPy, caller casts
    to fp32 (nerf.py:151). Ray n = y * W + x after reshape(-1, 3) (nerf.py:109)."""
    to_world = np.asarray(to_world, dtype=np.float64).reshape(4, 4)
    cam_loc = to_world[:3, 3]
    rayo = np.tile(cam_loc[None, None, :], (imh * sps, imw * sps, 1))
    xs = np.linspace(0, imw, imw * sps, endpoint=False)
    ys = np.linspace(0, imh, imh * sps, endpoint=False)
    xs, ys = np.meshgrid(xs, ys)
    fl = .5 * imw / np.tan(.5 * angle_x)
    rayd = np.stack(
        ((xs - .5 * imw) / fl, -(ys - .5 * imh) / fl, -np.ones_like(xs)), axis=-1)
    rayd = np.sum(rayd[:, :, np.newaxis, :] * to_world[:3, :3], axis=-1)
    return rayo.astype(np.float32), rayd.astype(np.float32)


# ------------------------------------------------------------------ sampling

def gen_z(near, far, n_samples, n_rays, lin_in_disp=False, perturb_u=None,
          dtype=torch.float32):
    """nerfactor/models/nerf.py:120-136. `perturb_u` [n_rays, n_samples] replaces
    tf.random.uniform (:134) when stratified perturbation is wanted."""
    t = tfops.linspace(0., 1., n_samples, dtype)
    if lin_in_disp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z[None, :].expand(n_rays, n_samples)
    if perturb_u is not None:
        mid = .5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * torch.as_tensor(perturb_u, dtype=dtype)
    return z


def inv_transform_sample(val, weights, n_samples, eps=1e-5):
    """nerfactor/util/math.py:71-94 with det=True."""
    dt = val.dtype
    denom = torch.sum(weights, -1, keepdim=True) + eps
    pdf = weights / denom
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat((torch.zeros_like(cdf[:, :1]), cdf), -1)
    u = tfops.linspace(0., 1., n_samples, dt)
    u = u[None, :].expand(cdf.shape[0], n_samples)
    ind = tfops.searchsorted_right(cdf, u)
    below = torch.clamp(ind - 1, min=0)
    above = torch.clamp(ind, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    val_b, val_a = torch.gather(val, 1, below), torch.gather(val, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return val_b + t * (val_a - val_b)


def gen_z_fine(z_coarse, weights, n_samples_fine):
    """nerfactor/models/nerf.py:138-147 (perturb=False)."""
    mid = .5 * (z_coarse[:, 1:] + z_coarse[:, :-1])
    z_fine = inv_transform_sample(mid, weights[..., 1:-1], n_samples_fine)
    z_all, _ = torch.sort(torch.cat((z_coarse, z_fine), -1), -1)
    return z_all


def accumulate_sigma(sigma, z, rayd, inf=1e10):
    """nerfactor/models/nerf.py:184-212 with noise_std=0."""
    dist = z[:, 1:] - z[:, :-1]
    dist = torch.cat(
        (dist, torch.full_like(dist[:, :1], inf)), dim=-1)
    dist = dist * torch.linalg.norm(rayd[:, None, :], dim=-1)
    density = 1.0 - torch.exp(-torch.relu(sigma) * dist)
    return density * tfops.safe_cumprod(1. - density)


# -------------------------------------------------------------- sigma network

def _sigma_raw(nerf, pts, use_fine, n_freqs_xyz=10):
    pref = 'fine_' if use_fine else 'coarse_'
    e = nets.embed(pts, n_freqs_xyz)
    return nets.mlp_forward(
        nerf[pref + 'sigma_out'], nets.mlp_forward(nerf[pref + 'enc'], e))


def eval_nerf_at(nerf, pts, views, use_fine=False, n_freqs_xyz=10, n_freqs_view=4):
    """nerfactor/models/nerf.py:254-290 (use_views = True): [P, 4] = (raw rgb, raw sigma);
    pts, views: [P, 3]."""
    pref = 'fine_' if use_fine else 'coarse_'
    feat = nets.mlp_forward(nerf[pref + 'enc'], nets.embed(pts, n_freqs_xyz))
    sigma = nets.mlp_forward(nerf[pref + 'sigma_out'], feat)
    feat = nets.mlp_forward(nerf[pref + 'bottleneck'], feat)
    feat_views = torch.cat((feat, nets.embed(views, n_freqs_view)), -1)
    rgb = nets.mlp_forward(nerf[pref + 'rgb_out'], feat_views)
    return torch.cat((rgb, sigma), -1)


def nerf_accumulate(rgbs, z, rayd, white_bg=True, eps=1e-10):
    """nerf.py:214-252: weights from sigma (accumulate_sigma), sigmoid colours, weighted sums,
    disparity, composite onto the background."""
    weights = accumulate_sigma(rgbs[:, :, 3], z, rayd)
    rgb = torch.sigmoid(rgbs[:, :, :3])
    occu = torch.sum(weights, dim=-1)
    rgb = torch.sum(weights[:, :, None] * rgb, dim=-2)
    depth = torch.sum(weights * z, dim=-1)
    disp = 1. / torch.clamp(depth, min=eps)
    bg = torch.ones_like(rgb) if white_bg else torch.zeros_like(rgb)
    rgb = rgb * occu[:, None] + bg * (1. - occu[:, None])          # imgutil.alpha_blend
    return rgb, occu, depth, disp, weights


def nerf_render_rays(nerf, rayo, rayd, near, far, n_samples_coarse=64, n_samples_fine=128,
                     lin_in_disp=False, white_bg=True):
    """nerf.py:149-182 at test time (perturb = False)."""
    rayd = l2_normalize(rayd, 1, 1e-12)
    n = rayo.shape[0]
    z = gen_z(near, far, n_samples_coarse, n, lin_in_disp, dtype=rayo.dtype)
    pts = rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]
    views = rayd[:, None, :].expand_as(pts)
    rgbs = eval_nerf_at(nerf, pts.reshape(-1, 3), views.reshape(-1, 3), False).reshape(n, -1, 4)
    rgb, occu, depth, disp, weights = nerf_accumulate(rgbs, z, rayd, white_bg)
    coarse = {'rgb': rgb, 'occu': occu, 'depth': depth, 'disp': disp}
    if n_samples_fine <= 0:
        return coarse, {}
    z = gen_z_fine(z, weights, n_samples_fine)
    pts = rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]
    views = rayd[:, None, :].expand_as(pts)
    rgbs = eval_nerf_at(nerf, pts.reshape(-1, 3), views.reshape(-1, 3), True).reshape(n, -1, 4)
    rgb, occu, depth, disp, _ = nerf_accumulate(rgbs, z, rayd, white_bg)
    return coarse, {'rgb': rgb, 'occu': occu, 'depth': depth, 'disp': disp}


def check_bounds(pts, scene_bbox=None):
    """geometry_from_nerf.py:365-378. scene_bbox = (x0,x1,y0,y1,z0,z1) or None."""
    if scene_bbox is None:
        return torch.ones((pts.shape[0],), dtype=torch.bool)
    x0, x1, y0, y1, z0, z1 = [float(v) for v in scene_bbox]
    return ((pts[:, 0] >= x0) & (pts[:, 0] <= x1) & (pts[:, 1] >= y0) &
            (pts[:, 1] <= y1) & (pts[:, 2] >= z0) & (pts[:, 2] <= z1))


def eval_sigma_mlp(nerf, pts, use_fine=False, scene_bbox=None, mlp_chunk=65536):
    """geometry_from_nerf.py:322-350: relu(sigma_out(enc(embed(p)))), 0 outside
    the bounding box."""
    in_b = check_bounds(pts, scene_bbox)
    pts_in = pts[in_b]
    chunks = []
    for i in range(0, pts_in.shape[0], mlp_chunk):
        chunks.append(torch.relu(_sigma_raw(nerf, pts_in[i:i + mlp_chunk], use_fine)))
    sigma = torch.zeros((pts.shape[0], 1), dtype=pts.dtype)
    if chunks:
        sigma[in_b] = torch.cat(chunks, 0)
    return sigma


def sigma_and_normal(nerf, pts, mlp_chunk=65536):
    """geometry_from_nerf.py:285-300: fine sigma and -l2_normalize(d sigma/dx)
    (batch_jacobian of the post-ReLU sigma w.r.t. raw xyz)."""
    sig, nrm = [], []
    for i in range(0, pts.shape[0], mlp_chunk):
        p = pts[i:i + mlp_chunk].detach().clone().requires_grad_(True)
        s = torch.relu(_sigma_raw(nerf, p, True))
        (g,) = torch.autograd.grad(s.sum(), p)
        sig.append(s.detach())
        nrm.append(-l2_normalize(g, 1))
    return torch.cat(sig, 0), torch.cat(nrm, 0)


# ---------------------------------------------------------- camera -> surface

def compute_depth_and_normal(nerf, rayo, rayd, near, far, n_samples_coarse=64,
                             n_samples_fine=128, lin_in_disp=False,
                             scene_bbox=None, mlp_chunk=65536):
    """geometry_from_nerf.py:249-319.  n_samples_* are the ini values; the
    reference adds 64 to each (:250-251)."""
    n_c, n_f = 64 + n_samples_coarse, 64 + n_samples_fine
    n_rays = rayo.shape[0]
    z = gen_z(near, far, n_c, n_rays, lin_in_disp, dtype=rayo.dtype)
    pts = rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]
    sigma = eval_sigma_mlp(nerf, pts.reshape(-1, 3), False, scene_bbox,
                           mlp_chunk).reshape(n_rays, -1)
    weights = accumulate_sigma(sigma, z, rayd)
    z = gen_z_fine(z, weights, n_f)
    pts = rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]
    pts_flat = pts.reshape(-1, 3)
    in_b = check_bounds(pts_flat, scene_bbox)
    sigma_flat, normal_flat = sigma_and_normal(nerf, pts_flat, mlp_chunk)
    sigma_flat = torch.where(in_b[:, None], sigma_flat, torch.zeros_like(sigma_flat))
    sigma = sigma_flat.reshape(n_rays, -1)
    normal = normal_flat.reshape(pts.shape)
    weights = accumulate_sigma(sigma, z, rayd)
    occu = torch.sum(weights, -1)
    exp_depth = torch.sum(weights * z, dim=-1)
    exp_normal = torch.sum(weights[:, :, None] * normal, dim=-2)
    return occu, exp_depth, exp_normal


def march_single_pass(nerf, rayo, rayd, near, far, n_samples, use_fine=False,
                      perturb_u=None, scene_bbox=None, mlp_chunk=65536):
    """The single-pass march the headline benchmark uses (SURVEY 8d 'Caveat on
    128 spp'): gen_z (nerf.py:120-136) -> eval_sigma_mlp
    (geometry_from_nerf.py:322-350) -> accumulate_sigma (nerf.py:184-212) ->
    occu = sum w, depth = sum w z (geometry_from_nerf.py:312-315) and
    surf = rayo + rayd * depth (:134)."""
    n_rays = rayo.shape[0]
    z = gen_z(near, far, n_samples, n_rays, perturb_u=perturb_u, dtype=rayo.dtype)
    pts = rayo[:, None, :] + rayd[:, None, :] * z[:, :, None]
    sigma = eval_sigma_mlp(nerf, pts.reshape(-1, 3), use_fine, scene_bbox,
                           mlp_chunk).reshape(n_rays, -1)
    weights = accumulate_sigma(sigma, z, rayd)
    occu = torch.sum(weights, -1)
    depth = torch.sum(weights * z, dim=-1)
    surf = rayo + rayd * depth[:, None]
    return {'z': z, 'sigma': sigma, 'weights': weights, 'occu': occu,
            'depth': depth, 'surf': surf}


# ----------------------------------------------------------- surface -> light

def compute_light_visibility(nerf, surf, normal, lxyz, lvis_far=1., lvis_near=.1,
                             n_samples_coarse=64, n_samples_fine=128,
                             lin_in_disp=False, scene_bbox=None, mlp_chunk=65536,
                             lpix_chunk=1):
    """geometry_from_nerf.py:177-246.  lxyz [L,3] fp32 (gen_light_xyz)."""
    n_c, n_f = 64 + n_samples_coarse, 64 + n_samples_fine
    dt = surf.dtype
    lxyz_flat = torch.as_tensor(np.asarray(lxyz, dtype=np.float32)).to(dt).reshape(1, -1, 3)
    n_lights = lxyz_flat.shape[1]
    lvis_hit = torch.zeros((surf.shape[0], n_lights), dtype=dt)
    for i in range(0, n_lights, lpix_chunk):
        end_i = min(n_lights, i + lpix_chunk)
        surf2l = lxyz_flat[:, i:end_i, :] - surf[:, None, :]
        surf2l = l2_normalize(surf2l, 2)
        surf2l_flat = surf2l.reshape(-1, 3)
        surf_flat = surf[:, None, :].expand(-1, surf2l.shape[1], -1).reshape(-1, 3)
        lcos = torch.einsum('ijk,ik->ij', surf2l, normal)
        front_lit = lcos > 0
        if front_lit.sum() == 0:
            continue
        fl = front_lit.reshape(-1)
        s_fl, d_fl = surf_flat[fl], surf2l_flat[fl]
        z = gen_z(lvis_near, lvis_far, n_c, d_fl.shape[0], lin_in_disp, dtype=dt)
        pts = s_fl[:, None, :] + d_fl[:, None, :] * z[:, :, None]
        sigma = eval_sigma_mlp(nerf, pts.reshape(-1, 3), False, scene_bbox,
                               mlp_chunk).reshape(pts.shape[:2])
        weights = accumulate_sigma(sigma, z, d_fl)
        z = gen_z_fine(z, weights, n_f)
        pts = s_fl[:, None, :] + d_fl[:, None, :] * z[:, :, None]
        sigma = eval_sigma_mlp(nerf, pts.reshape(-1, 3), True, scene_bbox,
                               mlp_chunk).reshape(pts.shape[:2])
        weights = accumulate_sigma(sigma, z, d_fl)
        occu = torch.sum(weights, -1)
        blk = lvis_hit[:, i:end_i]
        blk[front_lit] = 1 - occu
        lvis_hit[:, i:end_i] = blk
    return lvis_hit


# ------------------------------------------------------ process_view tail

def postprocess_view(occu, exp_depth, exp_normal, rayo, rayd, hw, occu_thres=0.):
    """geometry_from_nerf.py:122-149 for spp=1: alpha/xyz/normal maps."""
    h, w = hw
    occu = torch.where(occu < occu_thres, torch.zeros_like(occu), occu)
    alpha_map = torch.clamp(occu.reshape(h, w), 0., 1.)
    surf = rayo + rayd * exp_depth[:, None]
    xyz_map = surf.reshape(h, w, 3) * alpha_map[:, :, None]
    normal_map = exp_normal.reshape(h, w, 3)
    bg = torch.tensor((0., 1., 0.), dtype=occu.dtype)[None, None, :].expand(h, w, 3)
    a = alpha_map[:, :, None]
    normal_map = normal_map * a + bg * (1. - a)
    normal_map = l2_normalize(normal_map, 2)
    normal_map = torch.clamp(normal_map, -1., 1.)
    return alpha_map, xyz_map, normal_map, surf

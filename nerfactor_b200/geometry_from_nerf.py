"""Mirror of nerfactor/geometry_from_nerf.py (Stage A): camera -> surface march
(`compute_depth_and_normal`), surface -> light march (`compute_light_visibility`),
`eval_sigma_mlp`, plus `march_single_pass`, the single-pass S-sample march the
headline benchmark times.  Script-level flags of the reference
(`FLAGS.mlp_chunk`, `lpix_chunk`, `light_h`, `lvis_far`, `scene_bbox`,
`occu_thres`) are keyword arguments here."""
import numpy as np
import torch

from . import _lib
from .brdf.renderer import gen_light_xyz


def parse_bbox(scene_bbox):
    """geometry_from_nerf.py:365-371: 'x_min,x_max,y_min,y_max,z_min,z_max' or None."""
    if scene_bbox is None or scene_bbox == '':
        return None
    if isinstance(scene_bbox, str):
        return [float(v) for v in scene_bbox.split(',')]
    return [float(v) for v in scene_bbox]


def eval_sigma_mlp(model, rayo, rayd, z, use_fine=False, scene_bbox=None, precision=None):
    """geometry_from_nerf.py:322-350 on the samples o + z d (never materialised):
    relu(sigma_out(enc(embed(p)))), 0 outside the bounding box."""
    prec = precision or model.precision
    return _lib.sigma_fwd(model.ctx, model.packed_sigma(use_fine), rayo, rayd, z,
                          parse_bbox(scene_bbox), prec)


def march_single_pass(model, rayo, rayd, n_samples, use_fine=False, perturb_u=None,
                      scene_bbox=None, precision=None, want_weights=False):
    """gen_z (nerf.py:120-136) -> sigma (gfn.py:322-350) -> weights (nerf.py:184-212)
    -> occu / depth (gfn.py:312-315) -> surf = rayo + rayd * depth (gfn.py:134)."""
    ctx = model.ctx
    n = rayo.shape[0]
    z = _lib.gen_z(ctx, model.near, model.far, n_samples, n, False, perturb_u)
    sigma = eval_sigma_mlp(model, rayo, rayd, z, use_fine, scene_bbox, precision)
    w, occu, depth, surf, _ = _lib.composite(ctx, sigma, z, rayo, rayd,
                                             want_weights=want_weights)
    return {'z': z, 'sigma': sigma, 'weights': w, 'occu': occu, 'depth': depth,
            'surf': surf}


def compute_depth_and_normal(model, rayo, rayd, config, scene_bbox=None, precision=None):
    """geometry_from_nerf.py:249-319 -> (occu[N], exp_depth[N], exp_normal[N,3]): one call into
    nf_raymarch_depth_normal_fwd (coarse march, inverse-CDF resampling, fine march with
    d sigma / dx normals, compositing; all [N, S] intermediates in a workspace)."""
    n_c = 64 + config.getint('DEFAULT', 'n_samples_coarse')
    n_f = 64 + config.getint('DEFAULT', 'n_samples_fine')
    lin = config.getboolean('DEFAULT', 'lin_in_disp')
    near, far = config.getfloat('DEFAULT', 'near'), config.getfloat('DEFAULT', 'far')
    prec = precision or model.precision
    return _lib.raymarch_depth_normal_fwd(
        model.ctx, model.packed_sigma(False), model.packed_sigma(True), rayo, rayd, near, far,
        n_c, n_f, lin, parse_bbox(scene_bbox), prec)


def compute_light_visibility(model, surf, normal, config, lvis_near=.1, lvis_far=1.,
                             light_h=16, scene_bbox=None, precision=None, lxyz=None):
    """geometry_from_nerf.py:177-246 -> lvis_hit [M, L] (device tensor): one call into
    nf_raymarch_lvis_fwd.  All lights are marched together in chunks of (point, light) pairs
    instead of the reference's 512-iteration Python loop; back-lit pairs stay 0."""
    ctx = model.ctx
    n_c = 64 + config.getint('DEFAULT', 'n_samples_coarse')
    n_f = 64 + config.getint('DEFAULT', 'n_samples_fine')
    lin = config.getboolean('DEFAULT', 'lin_in_disp')
    if lxyz is None:
        lxyz, _ = gen_light_xyz(light_h, 2 * light_h)
    lxyz = torch.as_tensor(np.asarray(lxyz, np.float32).reshape(-1, 3)).to(ctx.device)
    return _lib.raymarch_lvis_fwd(
        ctx, model.packed_sigma(False), model.packed_sigma(True), surf.contiguous(),
        normal.contiguous(), lxyz, lvis_near, lvis_far, n_c, n_f, lin, parse_bbox(scene_bbox),
        precision or model.precision)


def postprocess_view(occu, exp_depth, exp_normal, rayo, rayd, hw, occu_thres=0.):
    """geometry_from_nerf.py:122-149 (spp = 1) on device tensors: alpha map (thresholded,
    clipped), alpha-premultiplied xyz map, normal map blended towards (0, 1, 0) and
    re-normalised.  Returns (alpha_map [H,W], xyz_map [H,W,3], normal_map [H,W,3], surf [N,3])."""
    h, w = hw
    occu = torch.where(occu < occu_thres, torch.zeros_like(occu), occu)
    alpha_map = torch.clamp(occu.reshape(h, w), 0., 1.)
    surf = rayo + rayd * exp_depth[:, None]
    a = alpha_map[:, :, None]
    xyz_map = surf.reshape(h, w, 3) * a
    bg = torch.tensor((0., 1., 0.), device=occu.device)[None, None, :]
    normal_map = exp_normal.reshape(h, w, 3) * a + bg * (1. - a)
    sq = torch.sum(normal_map * normal_map, dim=2, keepdim=True)
    normal_map = normal_map * torch.rsqrt(torch.clamp(sq, min=1e-12))
    return alpha_map, xyz_map, torch.clamp(normal_map, -1., 1.), surf


def process_view(model, rayo, rayd, hw, config, occu_thres=0., lvis_far=1., light_h=16,
                 scene_bbox=None, precision=None, with_lvis=True):
    """geometry_from_nerf.py:93-174 without the file I/O: the buffers the reference writes as
    alpha.png / xyz.npy / normal.npy / lvis.npy, returned as device tensors."""
    occu, depth, normal = compute_depth_and_normal(model, rayo, rayd, config, scene_bbox,
                                                   precision)
    alpha_map, xyz_map, normal_map, surf = postprocess_view(occu, depth, normal, rayo, rayd,
                                                            hw, occu_thres)
    out = {'alpha': alpha_map, 'xyz': xyz_map, 'normal': normal_map}
    if with_lvis:
        hit = alpha_map.reshape(-1) > 0.                                   # gfn.py:154
        idx = torch.nonzero(hit, as_tuple=False)[:, 0]
        lvis_hit = compute_light_visibility(
            model, surf.index_select(0, idx).contiguous(), normal.index_select(0, idx).contiguous(),
            config, lvis_far=lvis_far, light_h=light_h, scene_bbox=scene_bbox,
            precision=precision)
        lvis_hit = torch.clamp(lvis_hit, 0., 1.)                           # gfn.py:160
        L = lvis_hit.shape[1]
        lvis = torch.zeros((hw[0] * hw[1], L), device=surf.device)
        lvis.index_copy_(0, idx, lvis_hit)
        out['lvis'] = lvis.reshape(hw[0], hw[1], L) * alpha_map[:, :, None]   # gfn.py:170-171
    return out


# =============================================================================== script
def _parse_args(argv=None):
    """geometry_from_nerf.py:30-58 (same flag names; `mlp_chunk` / `lpix_chunk` are accepted and
    ignored: the fused kernels are persistent over tiles and march all lights together)."""
    import argparse
    ap = argparse.ArgumentParser(description="Stage A: surface, normals and light visibility "
                                             "from a trained NeRF")
    ap.add_argument('--trained_nerf', default='',
                    help="path to trained NeRF up to (and including) learning rate folder")
    ap.add_argument('--data_root', default='', help="input data root")
    ap.add_argument('--out_root', default='', help="output root")
    ap.add_argument('--imh', type=int, default=None,
                    help="image height (defaults to what was used for NeRF training)")
    ap.add_argument('--scene_bbox', default=None,
                    help="x_min,x_max,y_min,y_max,z_min,z_max")
    ap.add_argument('--lvis_far', type=float, default=1.)
    ap.add_argument('--occu_thres', type=float, default=0.)
    ap.add_argument('--light_h', type=int, default=16)
    ap.add_argument('--mlp_chunk', type=int, default=1_500_000)
    ap.add_argument('--lpix_chunk', type=int, default=1)
    ap.add_argument('--spp', type=int, default=1)
    ap.add_argument('--fps', type=int, default=12)
    ap.add_argument('--debug', action='store_true')
    ap.add_argument('--precision', default='f16e', choices=['f16e', 'f16', 'bf16', 'fp32'],
                    help="'fp32' = CUDA-core kernels throughout (tight parity), else tcgen05; "
                         "'f16e' = fp16 operands with the positional encoding as an fp16 hi + lo pair")
    return ap.parse_args(argv)


def main(argv=None):
    """geometry_from_nerf.py:63-90: latest NeRF checkpoint -> for every train / vali / test view
    the four geometry buffers under `<out_root>/<view>/` (views already done are skipped,
    :106-115).  Under torchrun the views are split round-robin over the ranks."""
    import os
    from os.path import basename, join
    from . import datasets, models
    from .util import config as configutil, geom_io, io as ioutil

    FLAGS = _parse_args(argv)
    if FLAGS.spp != 1:
        # the reference's visibility step masks un-averaged points with an [H*W] mask
        # (gfn.py:154-156): only spp = 1 is self-consistent there
        raise NotImplementedError("spp != 1")
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if torch.cuda.is_available():          # without a GPU the model constructor raises (no CPU path)
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    ckpts = ioutil.sortglob(join(FLAGS.trained_nerf, 'checkpoints'), 'ckpt-*', ext='index')
    assert ckpts, "no checkpoint under %s/checkpoints" % FLAGS.trained_nerf
    ckpt_ind = [int(basename(x)[len('ckpt-'):-len('.index')]) for x in ckpts]
    latest_ckpt = ckpts[int(np.argmax(ckpt_ind))][:-len('.index')]
    config = ioutil.read_config(configutil.get_config_ini(latest_ckpt))
    if FLAGS.imh is not None:
        config.set('DEFAULT', 'imh', str(FLAGS.imh))
    if FLAGS.data_root:
        config.set('DEFAULT', 'data_root', FLAGS.data_root)
    Model = models.get_model_class(config.get('DEFAULT', 'model', fallback='nerf'))
    model = Model(config, precision=FLAGS.precision)
    ioutil.restore_model(model, latest_ckpt)
    Dataset = datasets.get_dataset_class(config.get('DEFAULT', 'dataset', fallback='nerf'))
    done = []
    k = 0
    for mode in ('train', 'vali', 'test'):
        try:
            dataset = Dataset(config, mode, always_all_rays=True, spp=FLAGS.spp)
        except AssertionError:            # no view of this mode
            continue
        for batch in dataset.build_pipeline(no_batch=True, no_shuffle=True):
            k += 1
            if (k - 1) % world != rank:
                continue
            id_, hw, rayo, rayd, _ = batch
            out_dir = join(FLAGS.out_root, id_)
            if geom_io.view_done(out_dir):
                continue
            rayo = rayo.to(model.device, non_blocking=True)
            rayd = rayd.to(model.device, non_blocking=True)
            rayd = rayd * torch.rsqrt(torch.clamp((rayd * rayd).sum(1, keepdim=True), min=1e-12))
            with torch.no_grad():
                buffers = process_view(
                    model, rayo.contiguous(), rayd.contiguous(), hw, config,
                    occu_thres=FLAGS.occu_thres, lvis_far=FLAGS.lvis_far, light_h=FLAGS.light_h,
                    scene_bbox=FLAGS.scene_bbox, precision=FLAGS.precision)
            geom_io.write_view_buffers(buffers, out_dir)
            done.append(id_)
            if FLAGS.debug:
                break
    return done


if __name__ == '__main__':
    main()

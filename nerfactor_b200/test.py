"""Mirror of nerfactor/test.py: relighting / view-synthesis inference over the test views.

    python -m nerfactor_b200.test --ckpt <outroot>/<xname>/checkpoints/ckpt-N \\
        [--color_correct_albedo] [--tgt_albedo gold|aluminium|green|rainbow|turbo] \\
        [--tgt_brdf <name>] [--sv_axis_i 0 --sv_axis_min -1.5 --sv_axis_max 1.5] [--debug]

Same flags, same output tree (`<outroot>/<xname>/vis_test/ckpt-N[_<edit>]/batch?????????/` +
a compiled .mp4).  Every view runs `model.call(batch, 'test', relight_olat=<final view>,
relight_probes=True, ...)` (test.py:180-186) through the fused kernels.  Under `torchrun` the
test views are split round-robin over the ranks (views are independent; no collective).
"""
import argparse
import os
from os.path import basename, join

import numpy as np
import torch

from . import datasets, models
from .util import config as configutil, img as imgutil, io as ioutil

# Chebyshev fit (degree 14, on 2x - 1) of Google's public "turbo" colormap, which the reference
# reads from a 256-entry table (third_party/turbo_colormap); max deviation from the table 0.009.
_TURBO_CHEB = (
    (0.50518, 0.30924, -0.12045, -0.23280, -0.04687, 0.09638, -0.01629, -0.03595, 0.02400,
     0.00210, -0.01475, 0.00802, 0.00488, -0.00942, 0.00060),
    (0.43698, -0.09692, -0.46156, 0.06777, 0.08082, 0.00387, -0.01191, -0.00327, -0.00329,
     0.00014, 0.00197, 0.00108, 0.00101, -0.00050, -0.00318),
    (0.35428, -0.35924, -0.12060, 0.22567, -0.14588, 0.04138, 0.04649, -0.02316, -0.01786,
     0.00113, 0.01111, 0.00662, -0.00772, -0.00456, 0.00708))
_RAINBOW = [(0.58, 0, 0.83), (0.29, 0, 0.51), (0, 0, 1), (0, 1, 0), (1, 1, 0), (1, 0.5, 0),
            (1, 0, 0)]
_FLAT = {'aluminium': (0.913, 0.921, 0.925), 'gold': (1, 0.843, 0), 'green': (0, 1, 0)}


def turbo(x):
    """interpolate_or_clip(turbo_colormap_data, x): black below 0, white above 1."""
    from numpy.polynomial import chebyshev
    x = np.asarray(x, np.float64)
    rgb = np.stack([chebyshev.chebval(2 * np.clip(x, 0, 1) - 1, c) for c in _TURBO_CHEB], -1)
    rgb = np.clip(rgb, 0., 1.)
    rgb[x < 0] = 0.
    rgb[x > 1] = 1.
    return rgb


def get_albedo_override(xyz, tgt_albedo, sv_axis_i=0, sv_axis_min=-1.5, sv_axis_max=1.5):
    """test.py:91-132 -> (3,) or [N,3] float32 array."""
    if tgt_albedo in _FLAT:
        return np.asarray(_FLAT[tgt_albedo], np.float32)
    xyz = xyz.detach().cpu().numpy() if hasattr(xyz, 'detach') else np.asarray(xyz)
    axis = xyz[:, sv_axis_i]
    if tgt_albedo == 'rainbow':
        band_width = (sv_axis_max - sv_axis_min) / len(_RAINBOW)
        out = np.zeros_like(xyz, dtype=np.float32)
        for i, color in enumerate(_RAINBOW):
            in_band = (axis >= sv_axis_min + i * band_width) & \
                (axis < sv_axis_min + (i + 1) * band_width)
            out[in_band] = color
        return out
    if tgt_albedo == 'turbo':
        return turbo((axis - sv_axis_min) / (sv_axis_max - sv_axis_min)).astype(np.float32)
    raise NotImplementedError("Target albedo: %s" % tgt_albedo)


def compute_rgb_scales(ckpt, alpha_thres=0.9):
    """test.py:46-88: per-channel least-squares scales matching the predicted albedo of the
    first validation view (latest epoch) to the ground-truth albedo, in linear space."""
    config_ini = configutil.get_config_ini(ckpt)
    config = ioutil.read_config(config_ini)
    vali_dir = join(config_ini[:-4], 'vis_vali')
    data_root = config.get('DEFAULT', 'data_root')
    epoch_dir = ioutil.sortglob(vali_dir, 'epoch?????????')[-1]
    batch_dir = ioutil.sortglob(epoch_dir, 'batch?????????')[0]
    view = ioutil.read_json(join(batch_dir, 'metadata.json'))['id']
    pred = imgutil.normalize_uint(imgutil.read(join(batch_dir, 'pred_albedo.png')))
    gt = imgutil.normalize_uint(imgutil.read(join(data_root, view, 'albedo.png')))
    pred = pred[:, :, :3] ** 2.2                      # undo the gamma of vis_batch
    gt = imgutil.resize(gt, new_h=pred.shape[0])      # method='tf' in the reference
    alpha, gt = gt[:, :, 3], gt[:, :, :3]
    is_fg = alpha > alpha_thres
    scales = []
    for i in range(3):
        x_hat, x = pred[:, :, i][is_fg], gt[:, :, i][is_fg]
        scales.append(x_hat.dot(x) / x_hat.dot(x_hat))
    return np.asarray(scales, np.float32)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--ckpt', default='/path/to/ckpt-100', help="path to checkpoint (prefix only)")
    ap.add_argument('--color_correct_albedo', action='store_true')
    ap.add_argument('--sv_axis_i', type=int, default=0,
                    help="along which axis we do spatially-varying edits")
    ap.add_argument('--sv_axis_min', type=float, default=-1.5)
    ap.add_argument('--sv_axis_max', type=float, default=1.5)
    ap.add_argument('--tgt_albedo', default=None, help="albedo edit name")
    ap.add_argument('--tgt_brdf', default=None, help="BRDF edit name")
    ap.add_argument('--debug', action='store_true')
    ap.add_argument('--precision', default='f16', choices=['f16', 'bf16', 'fp32'])
    ap.add_argument('--no_video', action='store_true', help="skip the final .mp4 compilation")
    return ap.parse_args(argv)


def main(argv=None):
    FLAGS = parse_args(argv)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    config_ini = configutil.get_config_ini(FLAGS.ckpt)
    config = ioutil.read_config(config_ini)
    outroot = join(config_ini[:-4], 'vis_test', basename(FLAGS.ckpt))
    if FLAGS.tgt_albedo:
        outroot = outroot.rstrip('/') + '_%s' % FLAGS.tgt_albedo
    if FLAGS.tgt_brdf:
        outroot = outroot.rstrip('/') + '_%s' % FLAGS.tgt_brdf
    # dataset
    Dataset = datasets.get_dataset_class(config.get('DEFAULT', 'dataset'))
    dataset = Dataset(config, 'test', debug=FLAGS.debug)
    n_views = dataset.get_n_views()
    datapipe = dataset.build_pipeline(
        no_batch=config.getboolean('DEFAULT', 'no_batch', fallback=True), no_shuffle=True)
    # model
    Model = models.get_model_class(config.get('DEFAULT', 'model'))
    model = Model(config, debug=FLAGS.debug, precision=FLAGS.precision)
    ioutil.restore_model(model, FLAGS.ckpt)
    albedo_scales = None
    if (not FLAGS.tgt_albedo) and FLAGS.color_correct_albedo:
        albedo_scales = compute_rgb_scales(FLAGS.ckpt)
    brdf_z_override = None
    if FLAGS.tgt_brdf:                                        # test.py:168-172
        bm = model.brdf_model
        brdf_z_override = np.asarray(bm.latent_code.z)[bm.brdf_names.index(FLAGS.tgt_brdf), :]
    for batch_i, batch in enumerate(datapipe):
        relight_olat = batch_i == n_views - 1                # only for the final view
        if batch_i % world == rank:
            albedo_override = None
            if FLAGS.tgt_albedo:
                albedo_override = get_albedo_override(
                    batch[6], FLAGS.tgt_albedo, FLAGS.sv_axis_i, FLAGS.sv_axis_min,
                    FLAGS.sv_axis_max)
            with torch.no_grad():
                _, _, _, to_vis = model.call(
                    batch, mode='test', relight_olat=relight_olat, relight_probes=True,
                    albedo_scales=albedo_scales, albedo_override=albedo_override,
                    brdf_z_override=brdf_z_override)
            outdir = join(outroot, 'batch{i:09d}'.format(i=batch_i))
            model.vis_batch(to_vis, outdir, mode='test', olat_vis=relight_olat)
        if FLAGS.debug:
            break
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
        dist.barrier()
    view_at = None
    if rank == 0 and not FLAGS.no_video:
        batch_vis_dirs = ioutil.sortglob(outroot, 'batch?????????')
        view_at = model.compile_batch_vis(batch_vis_dirs, outroot, mode='test')
        print("Compilation available for viewing at\n\t%s" % view_at)
    return outroot, view_at


if __name__ == '__main__':
    main()

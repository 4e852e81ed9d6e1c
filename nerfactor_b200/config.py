"""Flat `[DEFAULT]` configs like the reference's ini files.

`default_config(name)` reproduces the hyper-parameters of nerfactor/config/
{nerfactor,nerfactor_microfacet,shape,brdf,nerf}.ini that the hot path reads;
dataset / checkpoint paths are left out (there is no dataset here)."""
from configparser import ConfigParser

# "Must-have" pipeline keys of every reference .ini (trainvali.py reads them); paths are left
# out (data_root, data_nerf_root, outroot, *_ckpt, test_envmap_dir are per site)
_PIPELINE = {
    'no_batch': 'True', 'cache': 'True', 'lr_decay_steps': '500_000', 'lr_decay_rate': '0.1',
    'clipnorm': '-1', 'clipvalue': '-1', 'vis_train_batches': '4', 'keep_recent_epochs': '-1',
    'overwrite': 'False', 'xname': 'lr{lr}', 'viewer_prefix': ''}

_COMMON_MLP = {
    'mlp_chunk': '65536', 'mlp_width': '128', 'mlp_depth': '4', 'mlp_skip_at': '2',
    'pos_enc': 'True', 'n_freqs_xyz': '10', 'n_freqs_ldir': '4', 'n_freqs_vdir': '4'}

_NERFACTOR = dict(_COMMON_MLP, **_PIPELINE, **{            # nerfactor/config/nerfactor.ini
    'dataset': 'nerf_shape', 'lr': '5e-3', 'epochs': '100', 'ckpt_period': '10',
    'vali_period': '10', 'vali_batches': '4', 'use_nerf_alpha': 'False',
    'model': 'nerfactor', 'loss': 'l2', 'imh': '512', 'light_h': '16', 'near': '2',
    'far': '6', 'ndc': 'False', 'white_bg': 'True', 'xyz_jitter_std': '0.01',
    'smooth_use_l1': 'True', 'shape_mode': 'finetune', 'normal_loss_weight': '0.1',
    'lvis_loss_weight': '0.1', 'normal_smooth_weight': '0.05',
    'lvis_smooth_weight': '0.05', 'albedo_slope': '0.77', 'albedo_bias': '0.03',
    'pred_brdf': 'True', 'albedo_smooth_weight': '0.05', 'brdf_smooth_weight': '0.01',
    'learned_brdf_scale': '1', 'light_init_max': '1', 'light_tv_weight': '5e-6',
    'light_achro_weight': '0', 'linear2srgb': 'True', 'n_rays_per_step': '1024',
    'olat_inten': '200', 'ambient_inten': '0'})

_MICROFACET = dict(_NERFACTOR, **{            # nerfactor/config/nerfactor_microfacet.ini
    'model': 'nerfactor_microfacet', 'rough_min': '0.1', 'default_rough': '0.3',
    'fresnel_f0': '0.04', 'brdf_smooth_weight': '0'})

_SHAPE = dict(_COMMON_MLP, **_PIPELINE, **{                # nerfactor/config/shape.ini
    'dataset': 'nerf_shape', 'lr': '1e-2', 'epochs': '200', 'ckpt_period': '100',
    'vali_period': '100', 'vali_batches': '4', 'imh': '512', 'near': '2', 'far': '6',
    'ndc': 'False', 'n_rays_per_step': '1024',
    'model': 'shape', 'loss': 'l2', 'light_h': '16', 'white_bg': 'True',
    'xyz_jitter_std': '0.01', 'smooth_use_l1': 'True', 'normal_loss_weight': '1',
    'lvis_loss_weight': '1'})     # no smoothness weights in shape.ini: shape.py:37-40 falls back to 0

_BRDF = {                                     # nerfactor/config/brdf.ini
    'model': 'brdf', 'dataset': 'brdf_merl', 'loss': 'l2', 'loss_transform': 'log', 'lr': '1e-2',
    'lr_decay_steps': '500_000', 'lr_decay_rate': '0.1', 'epochs': '50_000',
    'ckpt_period': '1_000', 'vali_period': '1_000', 'vali_batches': '4', 'n_rays_per_step': '1024',
    'clipnorm': '-1', 'clipvalue': '-1', 'keep_recent_epochs': '-1', 'overwrite': 'False',
    'xname': 'lr{lr}', 'cache': 'True', 'no_batch': 'True', 'viewer_prefix': '',
    'pos_enc': 'True', 'n_freqs': '2', 'z_dim': '3',
    'z_gauss_mean': '0.', 'z_gauss_std': '0.01', 'normalize_z': 'False',
    'mlp_chunk': '65536', 'mlp_width': '128', 'mlp_depth': '4', 'mlp_skip_at': '2'}

_NERF = dict(_PIPELINE, **{                   # nerfactor/config/nerf.ini
    'dataset': 'nerf', 'lr': '1e-4', 'epochs': '2_000', 'ckpt_period': '100',
    'vali_period': '100', 'vali_batches': '8', 'imh': '512', 'n_rays_per_step': '1024',
    'enc_skip_at': '4', 'enc_width': '256', 'act': 'relu',
    'model': 'nerf', 'loss': 'l2', 'near': '2', 'far': '6', 'ndc': 'False',
    'white_bg': 'True', 'lin_in_disp': 'False', 'perturb': 'True', 'noise_std': '0',
    'n_samples_coarse': '64', 'n_samples_fine': '128', 'use_views': 'True',
    'pos_enc': 'True', 'n_freqs_xyz': '10', 'n_freqs_view': '4', 'mlp_width': '256',
    'enc_depth': '8', 'mlp_chunk': '65536', 'accu_chunk': '65536'})

_ALL = {'nerfactor': _NERFACTOR, 'nerfactor_microfacet': _MICROFACET, 'shape': _SHAPE,
        'brdf': _BRDF, 'nerf': _NERF}


def default_config(name, **overrides):
    cfg = ConfigParser()
    for k, v in _ALL[name].items():
        cfg.set('DEFAULT', k, v)
    for k, v in overrides.items():
        cfg.set('DEFAULT', k, str(v))
    return cfg

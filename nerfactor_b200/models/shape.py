"""Mirror of nerfactor/models/shape.py (normal + light-visibility MLPs).

Same constructor / `call` / `compute_loss` surface and the same `net` /
`embedder` keys as the reference; the compute goes through the C ABI
(`nerfactor_b200._lib`) into the sm_100a kernels.
"""
import numpy as np
import torch

from .. import _lib
from ..brdf.renderer import gen_light_xyz
from ..networks import mlp
from ..networks.embedder import Embedder
from ..util import math as mathutil, img as imgutil
from .base import Model as BaseModel
from ._visualize import ShapeVis


def to_device(x, device, dtype=torch.float32):
    """Host array (ideally pinned) or tensor -> contiguous device tensor."""
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(
        device, non_blocking=True)


class Model(ShapeVis, BaseModel):
    def __init__(self, config, debug=False, params=None, ctx=None, precision='f16'):
        super().__init__(config, debug=debug)
        self.ctx = ctx or _lib.default_context()
        self.device = self.ctx.device
        self.precision = precision          # arithmetic of the per-(point,light) nets
        self.point_precision = 'f16x3'      # per-point nets (see _pred_point)
        self.white_bg = self.config.getboolean('DEFAULT', 'white_bg')
        self.mlp_chunk = self.config.getint('DEFAULT', 'mlp_chunk')   # unused: fused kernels
        self.normal_smooth_weight = self.config.getfloat(
            'DEFAULT', 'normal_smooth_weight', fallback=0.)
        self.lvis_smooth_weight = self.config.getfloat(
            'DEFAULT', 'lvis_smooth_weight', fallback=0.)
        self.embedder = self._init_embedder()
        self.net = self._init_net()
        self.xyz_scale = self.config.getfloat('DEFAULT', 'xyz_scale', fallback=1.)
        lxyz, lareas = self._gen_lights()
        self.lxyz, self.lareas = lxyz, lareas
        self._packed = {}
        self._build_nets()
        if params is not None:
            self.load_params(params)

    # ------------------------------------------------------------ construction
    def _gen_lights(self):
        """shape.py:59-77: the lat-long grid of brdf/renderer.py, or -- MVS geometry, whose scale
        and scene centre differ -- the light positions stored in `<mvs_root>/lights.npz`."""
        mvs_root = self.config.get('DEFAULT', 'mvs_root', fallback=None)
        if mvs_root is None:
            light_h = self.config.getint('DEFAULT', 'light_h')
            lxyz, lareas = gen_light_xyz(light_h, int(2 * light_h))
        else:
            import os
            with open(os.path.join(mvs_root, 'lights.npz'), 'rb') as h:
                data = dict(np.load(h))
            lxyz, lareas = data['lxyzs'], data['lareas']
        return (torch.as_tensor(np.asarray(lxyz, np.float32)).to(self.device),
                torch.as_tensor(np.asarray(lareas, np.float32)).to(self.device))

    def set_lights(self, lxyz, lareas):
        """Arbitrary light sets (the reference's mvs_root/lights.npz path,
        shape.py:67-74; also the L != 2 h^2 configs of SURVEY 8d)."""
        self.lxyz = to_device(np.asarray(lxyz, np.float32), self.device)
        self.lareas = to_device(np.asarray(lareas, np.float32), self.device)

    def _init_net(self):
        """shape.py:79-94."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'mlp_depth')
        s = self.config.getint('DEFAULT', 'mlp_skip_at')
        net = {}
        net['normal_mlp'] = mlp.Network([w] * d, act=['relu'] * d, skip_at=[s])
        net['normal_out'] = mlp.Network([3], act=None)
        net['lvis_mlp'] = mlp.Network([w] * d, act=['relu'] * d, skip_at=[s])
        net['lvis_out'] = mlp.Network([1], act=['sigmoid'])
        return net

    def _init_embedder(self):
        """shape.py:96-126."""
        if not self.config.getboolean('DEFAULT', 'pos_enc'):
            raise NotImplementedError("pos_enc=False (tf.identity embedders)")
        mk = lambda n: Embedder(incl_input=True, in_dims=3, log2_max_freq=n - 1, n_freqs=n,
                                log_sampling=True)
        return {'xyz': mk(self.config.getint('DEFAULT', 'n_freqs_xyz')),
                'ldir': mk(self.config.getint('DEFAULT', 'n_freqs_ldir')),
                'vdir': mk(self.config.getint('DEFAULT', 'n_freqs_vdir'))}

    def _net_in_dims(self):
        dx, dl = self.embedder['xyz'].out_dims, self.embedder['ldir'].out_dims
        return {'normal': dx, 'lvis': dx + dl, 'albedo': dx, 'brdf_z': dx}

    def _build_nets(self, rng=None):
        """Keras builds Dense layers lazily on first call; here once, up front."""
        rng = rng or np.random.default_rng(0)
        width = self.config.getint('DEFAULT', 'mlp_width')
        for name, in_dim in self._net_in_dims().items():
            if name + '_mlp' in self.net:
                self.net[name + '_mlp'].build(in_dim, rng)
                self.net[name + '_out'].build(width, rng)

    def load_params(self, params):
        """params: name -> {'layers': [(W, b), ...]} (synth.make_stage_b_params)."""
        for k, net in self.net.items():
            if k in params:
                net.load(params[k])
        self._packed.clear()

    def weights_changed(self):
        """Call after editing layer weights so they are re-packed for the GPU."""
        self._packed.clear()

    def _packed_mlp(self, name, kind, **kw):
        if name not in self._packed:
            trunk, head = self.net[name + '_mlp'], self.net[name + '_out']
            layers = trunk.weights() + head.weights()
            self._packed[name] = _lib.PackedMlp(
                self.ctx, kind, layers, trunk.skip_at[0], head.layers[0].activation, **kw)
        return self._packed[name]

    # ------------------------------------------------------------ geometry glue
    def _calc_ldir(self, pts):
        """shape.py:128-135.  Materialises [N,L,3]; the kernels never call this
        (they form each direction on chip), it exists for API parity."""
        surf2l = self.lxyz.reshape(1, -1, 3) - pts[:, None, :]
        return mathutil.safe_l2_normalize(surf2l, axis=2)

    @staticmethod
    def _calc_vdir(cam_loc, pts):
        """shape.py:137-144."""
        return mathutil.safe_l2_normalize(cam_loc - pts, axis=1)

    @staticmethod
    def chunk_apply(func, x, dim, chunk_size):
        """shape.py:184-194.  Chunking bounded TF's memory; the fused kernels are
        persistent over tiles, so this is a single call."""
        return func(x)

    # ------------------------------------------------------------ network evals
    def _pred_point(self, name, pts):
        """Per-point nets: fp32-accurate always -- FP32 CUDA cores when the model runs in
        'fp32', otherwise tensor cores with the 3-term fp16 split ('f16x3')."""
        m = self._packed_mlp(name, 'point', n_freqs_a=self.embedder['xyz'].n_freqs)
        prec = 'fp32' if self.precision == 'fp32' else self.point_precision
        try:
            return _lib.point_mlp_fwd(self.ctx, m, pts, self.xyz_scale, prec)
        except _lib.NfError as e:
            if prec != 'fp32' and 'NF_ERR_UNSUPPORTED' in str(e):
                # network shape without a tcgen05 kernel: the FP32 CUDA kernel (still the GPU)
                return _lib.point_mlp_fwd(self.ctx, m, pts, self.xyz_scale, 'fp32')
            raise

    def _pred_normal_at(self, pts, eps=1e-6):
        """shape.py:196-211."""
        return self._pred_point('normal', pts) + eps

    def _pred_lvis_at(self, pts, surf2l=None):
        """shape.py:213-237.  `surf2l` is accepted for signature parity and ignored:
        the kernel derives l2n(lxyz - pts) itself (shape.py:128-135).  The jittered evaluation,
        where the reference passes the UN-jittered point's directions, is
        `_pred_lvis_jitter_at`."""
        m = self._packed_mlp('lvis', 'lvis', n_freqs_a=self.embedder['xyz'].n_freqs,
                             n_freqs_b=self.embedder['ldir'].n_freqs)
        return _lib.lvis_fwd(self.ctx, m, pts, self.lxyz.reshape(-1, 3), self.xyz_scale,
                             self.precision)

    JITTER_EXACT_MAX_PAIRS = 1 << 22      # (point, light) pairs materialised for the jittered net

    def _pred_lvis_jitter_at(self, pts_jitter, pts):
        """The smoothness-loss evaluation of shape.py:170 / nerfactor.py:225: the visibility net at
        the JITTERED point but with the light directions of the un-jittered one (`surf2l` is
        computed once from `pts` in the reference).  Tensor-core precisions: the fused kernel with
        a separate direction origin (nf_lvis_dirs_fwd), exact semantics at any size.  'fp32': the
        net layer by layer on materialised [embed(x + noise) | embed(surf2l)] rows through the FP32
        Dense kernels (nf_dense_fwd) -- train batches are 1024 rays x 512 lights; above
        JITTER_EXACT_MAX_PAIRS pairs (full-view validation batches, whose NeRFactor loss ignores
        the jitter terms) the FP32 fused kernel, whose directions follow the jittered point
        (<= 1e-4 in the result)."""
        lxyz = self.lxyz.reshape(-1, 3)
        n, L = pts.shape[0], lxyz.shape[0]
        if n > 0 and self.precision != 'fp32':
            m = self._packed_mlp('lvis', 'lvis', n_freqs_a=self.embedder['xyz'].n_freqs,
                                 n_freqs_b=self.embedder['ldir'].n_freqs)
            return _lib.lvis_dirs_fwd(self.ctx, m, pts_jitter.contiguous(), pts.contiguous(), lxyz,
                                      self.xyz_scale, self.precision)
        if n == 0 or n * L > self.JITTER_EXACT_MAX_PAIRS:
            return self._pred_lvis_at(pts_jitter)
        from .. import autodiff as ad
        trunk, head = self.net['lvis_mlp'], self.net['lvis_out']
        if 'lvis_dev_layers' not in self._packed:
            self._packed['lvis_dev_layers'] = [
                (to_device(w, self.device), to_device(b, self.device))
                for w, b in trunk.weights() + head.weights()]
        acts = [l.activation for l in trunk.layers] + [l.activation for l in head.layers]
        with torch.no_grad():
            surf2l = mathutil.safe_l2_normalize(lxyz[None, :, :] - pts[:, None, :], axis=2)
            e_xyz = ad.embed(self.xyz_scale * pts_jitter, self.embedder['xyz'].n_freqs)
            e_dir = ad.embed(surf2l.reshape(-1, 3), self.embedder['ldir'].n_freqs)
            x = torch.cat((e_xyz[:, None, :].expand(n, L, e_xyz.shape[1]).reshape(n * L, -1),
                           e_dir), -1)
            out = ad.mlp_apply(x, self._packed['lvis_dev_layers'], acts, trunk.skip_at, 'fp32')
        return out.reshape(n, L)

    # ------------------------------------------------------------------- call
    def call(self, batch, mode='train', xyz_noise=None):
        """shape.py:146-182."""
        xyz_jitter_std = self.config.getfloat('DEFAULT', 'xyz_jitter_std')
        self._validate_mode(mode)
        id_, hw, _, _, _, alpha, xyz, normal, lvis = batch
        alpha, xyz, normal, lvis = [
            to_device(x, self.device) for x in (alpha, xyz, normal, lvis)]
        if xyz_noise is None and xyz_jitter_std > 0:
            xyz_noise = torch.randn_like(xyz) * xyz_jitter_std
        elif xyz_noise is not None:
            xyz_noise = to_device(xyz_noise, self.device)
        normal_pred = self._pred_normal_at(xyz)
        if xyz_noise is not None and self.normal_smooth_weight > 0:
            normal_jitter = self._pred_normal_at(xyz + xyz_noise)
        else:
            normal_jitter = None
        normal_pred = mathutil.safe_l2_normalize(normal_pred, axis=1)
        if normal_jitter is not None:
            normal_jitter = mathutil.safe_l2_normalize(normal_jitter, axis=1)
        lvis_pred = self._pred_lvis_at(xyz)
        if xyz_noise is not None and self.lvis_smooth_weight > 0:
            lvis_jitter = self._pred_lvis_jitter_at((xyz + xyz_noise).contiguous(), xyz)
        else:
            lvis_jitter = None
        pred = {'normal': normal_pred, 'lvis': lvis_pred}
        gt = {'normal': normal, 'lvis': lvis, 'alpha': alpha}
        loss_kwargs = {'normal_jitter': normal_jitter, 'lvis_jitter': lvis_jitter}
        to_vis = {'id': id_, 'hw': hw}
        for k, v in pred.items():
            to_vis['pred_' + k] = v
        for k, v in gt.items():
            to_vis['gt_' + k] = v
        return pred, gt, loss_kwargs, to_vis

    def compute_loss(self, pred, gt, **kwargs):
        """shape.py:239-277."""
        normal_loss_weight = self.config.getfloat('DEFAULT', 'normal_loss_weight')
        lvis_loss_weight = self.config.getfloat('DEFAULT', 'lvis_loss_weight')
        smooth_use_l1 = self.config.getboolean('DEFAULT', 'smooth_use_l1')
        mse = lambda a, b: torch.mean((a - b) ** 2, dim=-1)
        mae = lambda a, b: torch.mean(torch.abs(a - b), dim=-1)
        smooth_loss = mae if smooth_use_l1 else mse
        normal_jitter = kwargs.pop('normal_jitter')
        lvis_jitter = kwargs.pop('lvis_jitter')
        alpha = gt['alpha']
        bgv = 1. if self.white_bg else 0.
        blend = lambda x: imgutil.alpha_blend(x, alpha, torch.full_like(x, bgv))
        normal_pred, normal_gt = blend(pred['normal']), blend(gt['normal'])
        lvis_pred, lvis_gt = blend(pred['lvis']), blend(gt['lvis'])
        loss = normal_loss_weight * mse(normal_gt, normal_pred)
        loss = loss + lvis_loss_weight * mse(lvis_gt, lvis_pred)
        if normal_jitter is not None:
            loss = loss + self.normal_smooth_weight * smooth_loss(normal_pred, normal_jitter)
        if lvis_jitter is not None:
            loss = loss + self.lvis_smooth_weight * smooth_loss(lvis_pred, lvis_jitter)
        return loss

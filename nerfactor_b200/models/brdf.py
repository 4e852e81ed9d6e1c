"""Mirror of nerfactor/models/brdf.py: the MERL-prior BRDF MLP
(z ++ posenc(rusink) -> softplus scalar, brdf.py:57-66) and its latent codes.
The hot path needs the frozen forward (nerfactor.py:436-452, fused kernel nf_brdf_learned_fwd);
`call` / `compute_loss` (brdf.py:86-152) serve the prior's own training (trainvali.BrdfTrainer)
and evaluate the net on explicit Rusinkiewicz coordinates layer by layer through the Dense
kernels.  Its visualisation (brdf.py:154-329: matplotlib plots, MERL characteristic slices) is out
of scope."""
import numpy as np
import torch

from ..networks import mlp
from ..networks.embedder import Embedder
from ..networks.layers import LatentCode
from .base import Model as BaseModel


class Model(BaseModel):
    def __init__(self, config, debug=False, params=None, brdf_names=None, ctx=None,
                 precision=None):
        """`brdf_names`: the materials of the prior; by default read off `<data_root>/train_*.npz`
        like the reference (brdf.py:42-46), else four placeholder names."""
        super().__init__(config, debug=debug)
        if brdf_names is None:
            import os
            from ..util import io as ioutil
            root = self.config.get('DEFAULT', 'data_root', fallback='')
            found = ioutil.sortglob(root, 'train_*', ext='npz') if root and os.path.isdir(root) \
                else []
            brdf_names = [os.path.basename(x)[len('train_'):-len('.npz')] for x in found] or None
        self.mlp_chunk = self.config.getint('DEFAULT', 'mlp_chunk')
        self.embedder = self._init_embedder()
        self.net = self._init_net()
        self.brdf_names = list(brdf_names or ['synthetic_%03d' % i for i in range(4)])
        z_dim = self.config.getint('DEFAULT', 'z_dim')
        self.latent_code = LatentCode(
            len(self.brdf_names), z_dim,
            mean=self.config.getfloat('DEFAULT', 'z_gauss_mean'),
            std=self.config.getfloat('DEFAULT', 'z_gauss_std'),
            normalize=self.config.getboolean('DEFAULT', 'normalize_z'),
            rng=np.random.default_rng(0))
        width = self.config.getint('DEFAULT', 'mlp_width')
        rng = np.random.default_rng(1)
        self.net['brdf_mlp'].build(z_dim + self.embedder['rusink'].out_dims, rng)
        self.net['brdf_out'].build(width, rng)
        if params is not None:
            for k in ('brdf_mlp', 'brdf_out'):
                self.net[k].load(params[k])
        self.trainable = False

    def _init_net(self):
        """brdf.py:57-66."""
        w = self.config.getint('DEFAULT', 'mlp_width')
        d = self.config.getint('DEFAULT', 'mlp_depth')
        s = self.config.getint('DEFAULT', 'mlp_skip_at')
        return {'brdf_mlp': mlp.Network([w] * d, act=['relu'] * d, skip_at=[s]),
                'brdf_out': mlp.Network([1], act=['softplus'])}

    def _init_embedder(self):
        """brdf.py:68-85."""
        n = self.config.getint('DEFAULT', 'n_freqs')
        return {'rusink': Embedder(incl_input=True, in_dims=3, log2_max_freq=n - 1,
                                   n_freqs=n, log_sampling=True)}

    # ------------------------------------------------------------------ the prior's own forward
    def _device_layers(self):
        from .. import _lib
        from .shape import to_device
        dev = _lib.default_context().device
        return [(to_device(w, dev), to_device(b, dev)) for w, b in
                self.net['brdf_mlp'].weights() + self.net['brdf_out'].weights()], dev

    def _eval_brdf_at(self, z, rusink, layers=None, prec='fp32'):
        """brdf.py:113-138: softplus MLP on [z | embed(rusink)] and on the reciprocal coordinates
        (phi_d + pi), chunked like the reference (the chunking does not change values)."""
        from .. import autodiff as ad
        trunk, head = self.net['brdf_mlp'], self.net['brdf_out']
        acts = [l.activation for l in trunk.layers] + [l.activation for l in head.layers]
        n_freqs = self.embedder['rusink'].n_freqs
        shift = torch.tensor([np.pi, 0., 0.], dtype=torch.float32, device=rusink.device)
        out, out_reci = [], []
        for i in range(0, rusink.shape[0], self.mlp_chunk):
            zc, rc = z[i:i + self.mlp_chunk], rusink[i:i + self.mlp_chunk]
            for dst, r in ((out, rc), (out_reci, rc + shift)):
                x = torch.cat((zc, ad.embed(r, n_freqs)), 1)
                dst.append(ad.mlp_apply(x, layers, acts, trunk.skip_at, prec))
        return torch.cat(out, 0), torch.cat(out_reci, 0)

    def call(self, batch, mode='train'):
        """brdf.py:86-111.  batch = (id_, i, envmap_h, ims, spp, rusink [R,3], refl [R,1]); `i` is
        the material index (per batch or per row), -1 with an interpolation recipe in `id_`."""
        self._validate_mode(mode)
        id_, i, envmap_h, ims, spp, rusink, refl = batch
        layers, dev = self._device_layers()
        from .shape import to_device
        rusink, refl = to_device(rusink, dev), to_device(refl, dev)
        i0 = int(np.asarray(i).reshape(-1)[0])
        if mode == 'test' and i0 == -1:                       # novel identity: interpolate
            _, w1, mat1, w2, mat2 = str(id_).split('_')
            z = self.latent_code.interp(float(w1), self.brdf_names.index(mat1), float(w2),
                                        self.brdf_names.index(mat2))
        else:
            z = self.latent_code(i0)
        z = to_device(np.asarray(z, np.float32).reshape(1, -1), dev).expand(rusink.shape[0], -1)
        with torch.no_grad():
            brdf, brdf_reci = self._eval_brdf_at(z.contiguous(), rusink, layers)
        pred = {'brdf': brdf, 'brdf_reci': brdf_reci}
        gt = {'brdf': refl}
        to_vis = {'id': id_, 'i': i, 'z': z, 'gt_brdf': refl, 'envmap_h': envmap_h, 'ims': ims,
                  'spp': spp, **pred}
        return pred, gt, {}, to_vis

    def compute_loss(self, pred, gt, **kwargs):
        """brdf.py:140-152 with `loss = l2`: both the direct and the reciprocal prediction against
        the same ground truth, after `loss_transform` (log in brdf.ini)."""
        tr = self.config.get('DEFAULT', 'loss_transform', fallback='log')
        if tr.lower() == 'none':
            f = lambda x: x
        elif tr == 'log':
            f = torch.log
        elif tr == 'divide':
            f = lambda x: x / (x + 1.)
        else:
            raise NotImplementedError(tr)
        keep_batch = kwargs.get('keep_batch', False)
        loss = 0
        for p in (pred['brdf'], pred['brdf_reci']):
            per_row = torch.mean((f(p) - f(gt['brdf'])) ** 2, dim=-1)
            loss = loss + (per_row if keep_batch else per_row.mean())
        return loss

    def weights_changed(self):
        pass

    def vis_batch(self, data_dict, outdir, mode='train', dump_raw_to=None, n_vis=64):
        """brdf.py:154-329 reduced to its data: the reference draws bar plots (matplotlib) and, at
        test time, renders MERL characteristic slices; here the numbers behind those plots are
        written (`metadata.json`, `z.npy`, `log10_brdf.npy`: columns reciprocal prediction,
        prediction[, ground truth], a subset of `n_vis` rows)."""
        import os
        from ..util import io as ioutil
        self._validate_mode(mode)
        if mode == 'train':
            return
        os.makedirs(outdir, exist_ok=True)
        ioutil.write_json({'id': str(data_dict['id'])}, os.path.join(outdir, 'metadata.json'))
        tonp = lambda t: t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)
        np.save(os.path.join(outdir, 'z.npy'), tonp(data_dict['z'])[0])
        cols = [tonp(data_dict['brdf_reci']), tonp(data_dict['brdf'])]
        if mode == 'vali':
            cols.append(tonp(data_dict['gt_brdf']))
        val = np.hstack(cols)
        val = val[::max(1, int(val.shape[0] / n_vis)), :]
        np.save(os.path.join(outdir, 'log10_brdf.npy'), np.log10(val))

    def compile_batch_vis(self, batch_vis_dirs, outpref, mode='train'):
        """brdf.py: an HTML of the plots for validation; nothing to compile here."""
        self._validate_mode(mode)
        return None

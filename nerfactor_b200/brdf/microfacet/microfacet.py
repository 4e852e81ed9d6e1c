"""Mirror of brdf/microfacet/microfacet.py:21-111.

Inside the models the GGX lobe is fused into the rendering-equation kernel
(csrc/nf_integrate.cu: nothing of size [N, L, 3] is built); there this class only carries the
parameters `Model._eval_brdf_at` hands to `_render`.  Called directly -- the reference's
`Microfacet(...)(pts2l, pts2c, normal, albedo, rough)` -- it runs the standalone kernel
`nf_microfacet_brdf_fwd` on CUDA tensors and returns brdf[N, L, 3]."""
import numpy as np
import torch

from ... import _lib


class Microfacet:
    def __init__(self, default_rough=0.3, lambert_only=False, f0=0.91):
        self.default_rough = default_rough
        self.lambert_only = lambert_only
        self.f0 = f0

    def __call__(self, pts2l, pts2c, normal, albedo=None, rough=None, ctx=None):
        """pts2l [N, L, 3], pts2c [N, 3], normal [N, 3], albedo [N, 3] (None: ones), rough [N, 1]
        (None: default_rough), all in world coordinates (microfacet.py:30-72) -> [N, L, 3].
        NumPy inputs are uploaded; the result is a CUDA tensor.  No CPU path."""
        ctx = ctx or _lib.default_context()

        def dev(x):
            if x is None:
                return None
            t = torch.as_tensor(np.asarray(x, np.float32)) if not torch.is_tensor(x) else x
            return t.to(ctx.device, torch.float32).contiguous()
        pts2l, pts2c, normal, albedo, rough = [dev(x) for x in (pts2l, pts2c, normal, albedo, rough)]
        if pts2l.dim() != 3 or pts2l.shape[2] != 3 or pts2c.shape != (pts2l.shape[0], 3) or \
                normal.shape != pts2c.shape:
            raise ValueError("pts2l [N, L, 3], pts2c [N, 3], normal [N, 3] expected")
        if rough is not None:
            rough = rough.reshape(-1)
        return _lib.microfacet_brdf_fwd(ctx, pts2l, pts2c, normal, albedo, rough,
                                        self.default_rough, self.lambert_only, self.f0)

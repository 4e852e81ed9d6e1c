"""Mirror of brdf/microfacet/microfacet.py:21-72.  The GGX evaluation is fused
into the rendering-equation kernel (csrc/nf_integrate.cu); this class carries
the parameters and is what `Model._eval_brdf_at` hands to `_render`."""


class Microfacet:
    def __init__(self, default_rough=0.3, lambert_only=False, f0=0.91):
        if lambert_only:
            raise NotImplementedError("lambert_only=True has no caller on the hot path")
        self.default_rough = default_rough
        self.lambert_only = lambert_only
        self.f0 = f0

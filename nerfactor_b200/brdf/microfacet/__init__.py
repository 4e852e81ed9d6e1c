"""Mirror of the reference's `brdf/microfacet/` package: `brdf.microfacet.microfacet.Microfacet`
(brdf/microfacet/microfacet.py:21-111)."""
from .microfacet import Microfacet  # noqa: F401

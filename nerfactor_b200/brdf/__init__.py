"""Mirror of the reference's top-level `brdf` package (renderer, microfacet)."""
from . import renderer, microfacet  # noqa: F401

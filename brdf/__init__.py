"""Drop-in import root for the reference's top-level `brdf` package: `brdf.renderer`,
`brdf.microfacet.microfacet` resolve to `nerfactor_b200.brdf.*` (same module objects)."""
_nf_stub = True
from nerfactor_b200 import _aliases as _aliases  # noqa: E402

_aliases.install()

"""Mirror of nerfactor/networks (mlp.Network, Embedder, LatentCode)."""
from . import mlp, embedder, layers  # noqa: F401

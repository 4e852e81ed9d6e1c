"""Mirror of nerfactor/networks/seq.py:24-38: layers applied one after the other."""
from .base import Network as BaseNetwork
from . import mlp as _mlp


class Network(BaseNetwork):
    def build(self, input_shape, rng=None):
        import numpy as np
        rng = rng or np.random.default_rng()
        d = int(input_shape[-1]) if hasattr(input_shape, '__len__') else int(input_shape)
        for layer in self.layers:
            layer.build(d, rng)
            d = layer.units
        assert all(l.built for l in self.layers), "Some layers not built"

    def __call__(self, tensor):
        return _mlp.apply_layers(self.layers, tensor, skip_at=None)

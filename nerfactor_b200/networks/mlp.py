"""Mirror of nerfactor/networks/mlp.py:24-50 and seq.py:24-38.

`Network` holds the Dense layers of one trunk (or one output head) exactly like
the reference (`.layers`, in checkpoint order; `skip_at`).  The forward pass of
the hot path does not go through `Network.__call__`: `Model._pred_*_at` hand the
trunk + head pair to one fused CUDA kernel (embedding + Dense chain + head), so
`__call__` on a bare Network raises instead of silently running elsewhere.
"""
import math

import numpy as np


class Dense:
    """tf.keras.layers.Dense(w, activation=a): act(x @ kernel[in,out] + bias);
    glorot-uniform kernel, zero bias (mlp.py:34)."""

    def __init__(self, units, activation=None):
        self.units = units
        self.activation = activation
        self.kernel = None
        self.bias = None
        self.trainable = True

    @property
    def built(self):
        return self.kernel is not None

    def build(self, in_dim, rng):
        lim = math.sqrt(6.0 / (in_dim + self.units))
        self.kernel = rng.uniform(-lim, lim, size=(in_dim, self.units)).astype(np.float32)
        self.bias = np.zeros((self.units,), np.float32)

    def set_weights(self, kernel, bias):
        kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        assert kernel.ndim == 2 and kernel.shape[1] == self.units == bias.shape[0]
        self.kernel, self.bias = kernel, bias


class Network:
    def __init__(self, widths, act=None, skip_at=None):
        depth = len(widths)
        if act is None:
            act = [None] * depth
        assert len(act) == depth, \
            "If not `None`, `act` must have the save length as `widths`"
        self.layers = [Dense(w, a) for w, a in zip(widths, act)]
        self.skip_at = skip_at

    def build(self, in_dim, rng=None):
        """Builds every layer with the input size Keras would infer on first call
        (mlp.py:39-50: the layer after a skip sees width + in_dim)."""
        rng = rng or np.random.default_rng()
        d = in_dim
        for i, layer in enumerate(self.layers):
            layer.build(d, rng)
            d = layer.units
            if self.skip_at is not None and i in self.skip_at:
                d = layer.units + in_dim
        return self

    def load(self, mlp_dict):
        """mlp_dict: {'layers': [(W, b), ...]} (nerfactor_b200.synth / checkpoints)."""
        assert len(mlp_dict['layers']) == len(self.layers)
        for layer, (w, b) in zip(self.layers, mlp_dict['layers']):
            layer.set_weights(w, b)
        return self

    def weights(self):
        assert all(l.built for l in self.layers), "Some layers not built"
        return [(l.kernel, l.bias) for l in self.layers]

    def __call__(self, x):
        raise NotImplementedError(
            "nerfactor_b200 evaluates a trunk and its head inside one fused CUDA "
            "kernel (see Model._pred_*_at / nerfactor_b200._lib); calling a bare "
            "Network is not part of the hot path and has no fallback")

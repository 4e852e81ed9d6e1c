"""Mirror of nerfactor/networks/mlp.py:24-50 and seq.py:24-38.

`Network` holds the Dense layers of one trunk (or one output head) exactly like
the reference (`.layers`, in checkpoint order; `skip_at`).  The forward pass of
the hot path does not go through `Network.__call__`: `Model._pred_*_at` hand the
trunk + head pair to one fused CUDA kernel (embedding + Dense chain + head).
Calling a Network directly (`net(x)`, the reference surface, mlp.py:39-50) evaluates it
layer by layer on the FP32 Dense kernels (`nf_dense_fwd`) for a CUDA tensor `x`; there is no
CPU path, a CPU tensor raises.
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


from .base import Network as BaseNetwork


class Network(BaseNetwork):
    def __init__(self, widths, act=None, skip_at=None):
        super().__init__()
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

    def __call__(self, x):
        """x [M, in] (CUDA, fp32) -> [M, widths[-1]], input re-concatenated after the layers in
        `skip_at` as (y, x) (mlp.py:39-50)."""
        return apply_layers(self.layers, x, self.skip_at)


def apply_layers(layers, x, skip_at=None):
    """Dense chain on the FP32 Dense kernels (nf_dense_fwd through autodiff.mlp_apply)."""
    import torch
    from .. import autodiff as ad
    if not torch.is_tensor(x) or not x.is_cuda:
        raise TypeError("Network.__call__ needs a CUDA tensor: nerfactor_b200 has no CPU path "
                        "(the models call the fused kernels instead, see Model._pred_*_at)")
    assert all(l.built for l in layers), "Some layers not built"
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
    params = [(torch.as_tensor(l.kernel).to(x.device), torch.as_tensor(l.bias).to(x.device))
              for l in layers]
    y = ad.mlp_apply(x2, params, [l.activation for l in layers], skip_at, 'fp32')
    return y.reshape(*lead, y.shape[-1])

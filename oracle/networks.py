"""Oracle restatement of nerfactor/networks (test infrastructure).

Weights are plain dicts so the same parameter set feeds this oracle and the
CUDA product path:

    mlp = {'layers': [(W[in,out], b[out]), ...], 'act': [...], 'skip_at': [..]|None}
"""
import math

import numpy as np
import torch

from . import tfops


def embed(x, n_freqs, incl_input=True):
    """nerfactor/networks/embedder.py:23-47 with log_sampling=True and
    periodic_func=[sin, cos]: concat([x, sin(x*2^0), cos(x*2^0), sin(x*2^1), ...]);
    freq_bands = 2 ** linspace(0, n_freqs-1, n_freqs) (exact powers of two)."""
    out = []
    if incl_input:
        out.append(x)
    for k in range(n_freqs):
        freq = float(2.0 ** k)
        out.append(torch.sin(x * freq))
        out.append(torch.cos(x * freq))
    return torch.cat(out, dim=-1)


def embed_out_dims(in_dims, n_freqs, incl_input=True):
    return in_dims * ((1 if incl_input else 0) + 2 * n_freqs)


def _act(name, y):
    if name is None:
        return y
    if name == 'relu':
        return torch.relu(y)
    if name == 'sigmoid':
        return torch.sigmoid(y)
    if name == 'softplus':
        return torch.nn.functional.softplus(y)  # log(1 + e^y), keras 'softplus'
    raise NotImplementedError(name)


def mlp_forward(mlp, x):
    """nerfactor/networks/mlp.py:39-50 (skip) and seq.py:33-38 (sequential).
    Dense = act(x @ W[in,out] + b); the input is concatenated AFTER layer i for
    i in skip_at, as (y, x)."""
    layers, act, skip_at = mlp['layers'], mlp['act'], mlp.get('skip_at')
    dt = x.dtype
    x_ = x
    y = x
    for i, (w, b) in enumerate(layers):
        w_t = torch.as_tensor(w, dtype=dt)
        b_t = torch.as_tensor(b, dtype=dt)
        y = _act(act[i], x_ @ w_t + b_t)
        if skip_at is not None and i in skip_at:
            y = torch.cat((y, x), dim=-1)
        x_ = y
    return y


def glorot_uniform(rng, fan_in, fan_out):
    """Keras Dense default kernel init (networks/mlp.py:34): U(+-sqrt(6/(in+out)))."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def init_mlp(rng, in_dim, widths, act, skip_at=None, bias_std=0.0):
    """Builds a random-init MLP parameter dict with the layer input sizes the
    reference's lazily-built Keras Dense layers would get (mlp.py:39-50).
    bias_std > 0 perturbs the (Keras-zero) biases so bias handling is tested."""
    layers = []
    d = in_dim
    for i, w in enumerate(widths):
        W = glorot_uniform(rng, d, w)
        b = (rng.standard_normal(w) * bias_std).astype(np.float32)
        layers.append((W, b))
        d = w
        if skip_at is not None and i in skip_at:
            d = w + in_dim
    if act is None:
        act = [None] * len(widths)
    return {'layers': layers, 'act': list(act), 'skip_at': skip_at}

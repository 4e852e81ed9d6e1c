"""Host logic of the whole-network train calls on the CPU test double (tests/cpu_backend.py): the
weight padding / skip-layer splitting `autodiff.mlp_apply` does before `MlpChainFn` must give the
same values and gradients as the layer-by-layer path (`DenseFn` per layer)."""
import numpy as np
import pytest
import torch

import cpu_backend


@pytest.mark.parametrize('case', [(37, 90, (128, 128, 128, 128, 1), (2,), True),
                                  (20, 63, (128, 128, 3), None, False),
                                  (9, 18, (64, 128, 16, 4), (0,), True)])
def test_mlp_apply_chain_equals_layer_by_layer_on_test_double(monkeypatch, case):
    cpu_backend.install(monkeypatch)
    from nerfactor_b200 import autodiff as ad
    rows, in_dim, widths, skip_at, need_dx = case
    g = torch.Generator().manual_seed(rows)
    x = torch.randn((rows, in_dim), generator=g).requires_grad_(need_dx)
    layers, k = [], in_dim
    for i, n in enumerate(widths):
        kin = k + (in_dim if (skip_at and i - 1 in skip_at) else 0)
        layers.append(((torch.randn((kin, n), generator=g) / np.sqrt(kin)).requires_grad_(True),
                       (torch.randn((n,), generator=g) * 0.1).requires_grad_(True)))
        k = n
    acts = ['relu'] * (len(widths) - 1) + ['sigmoid']
    dy = torch.randn((rows, widths[-1]), generator=g)
    outs = {}
    for chain in (True, False):
        monkeypatch.setattr(ad, 'CHAIN', chain)
        y = ad.mlp_apply(x, layers, acts, skip_at, 'bf16')
        ins = [t for wb in layers for t in wb] + ([x] if need_dx else [])
        outs[chain] = (y.detach().clone(), [t.clone() for t in torch.autograd.grad(y, ins, dy)])
    assert outs[True][0].shape == (rows, widths[-1])
    assert torch.allclose(outs[True][0], outs[False][0], atol=1e-6)
    for a, b in zip(outs[True][1], outs[False][1]):
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-5, rtol=1e-5)

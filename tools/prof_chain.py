"""Profiling target: one nf_mlp_chain_fwd + _bwd on a visibility-network-sized batch
(524 288 rows, 92 -> 4 x 128 -> 4, skip after layer 2).
    ncu --set full -k regex:rowgemm_tc_kernel -c 2 python tools/prof_chain.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib          # noqa: E402

ctx = _lib.default_context()
rows, in_dim, widths, skip_layer = 524288, 92, (128, 128, 128, 128, 4), 3
g = torch.Generator(device='cpu').manual_seed(0)
x = torch.randn((rows, in_dim), generator=g).cuda()
ws, bs, k = [], [], in_dim
for i, n in enumerate(widths):
    kin = k + (in_dim if i == skip_layer else 0)
    ws.append((torch.randn((kin, n), generator=g) / kin ** .5).cuda())
    bs.append(torch.zeros((n,)).cuda())
    k = n
acts = ['relu'] * 4 + ['sigmoid']
for _ in range(2):
    y, work = _lib.mlp_chain_fwd(ctx, x, ws, bs, acts, skip_layer, 'bf16')
    dx, dws, dbs = _lib.mlp_chain_bwd(ctx, ws, bs, acts, skip_layer, in_dim, y, torch.ones_like(y), work,
                                      False, 'bf16')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y, work = _lib.mlp_chain_fwd(ctx, x, ws, bs, acts, skip_layer, 'bf16')
e1.record()
torch.cuda.synchronize()
print('chain fwd %.3f ms' % (e0.elapsed_time(e1) / 5))
e0.record()
for _ in range(5):
    _lib.mlp_chain_bwd(ctx, ws, bs, acts, skip_layer, in_dim, y, torch.ones_like(y), work, False, 'bf16')
e1.record()
torch.cuda.synchronize()
print('chain bwd %.3f ms' % (e0.elapsed_time(e1) / 5))

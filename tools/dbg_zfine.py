import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib
from oracle import stage_a
ctx = _lib.default_context()
rng = np.random.default_rng(0)
n, sc, sf = 257, 40, 56
u = rng.uniform(size=(n, sc)).astype(np.float32)
z_o = stage_a.gen_z(2., 6., sc, n, perturb_u=u)
z_g = _lib.gen_z(ctx, 2., 6., sc, n, False, torch.tensor(u).cuda())
w = rng.uniform(size=(n, sc)).astype(np.float32) ** 4
w[:5] = 0.
zf_o = stage_a.gen_z_fine(z_o, torch.tensor(w), sf).numpy()
zf_g = _lib.gen_z_fine(ctx, z_g, torch.tensor(w).cuda(), sf).cpu().numpy()
d = np.abs(zf_g - zf_o)
i = np.unravel_index(d.argmax(), d.shape)
print('maxdiff', d.max(), 'at', i, zf_g[i], zf_o[i], 'count>2e-5', (d > 2e-5).sum())
r = i[0]
print('row', r, 'got', zf_g[r][max(0,i[1]-3):i[1]+4], 'exp', zf_o[r][max(0,i[1]-3):i[1]+4])
rows = np.unique(np.nonzero(d > 2e-5)[0]); print('bad rows', rows[:20])

import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib
ctx = _lib.default_context()
rng = np.random.default_rng(0)
for K in (16, 64, 128):
    a = rng.integers(-4, 5, (256, K)).astype(np.float32)
    b = rng.integers(-4, 5, (128, K)).astype(np.float32)
    got = _lib.selftest_umma2(ctx, torch.tensor(a).cuda(), torch.tensor(b).cuda())
    torch.cuda.synchronize()
    got = got.cpu().numpy(); ref = a @ b.T
    print('K', K, 'maxerr', float(np.abs(got - ref).max()))
    if np.abs(got - ref).max() != 0:
        for rb in range(2):
            for cb in range(2):
                blk = np.abs(got[rb*128:(rb+1)*128, cb*64:(cb+1)*64] - ref[rb*128:(rb+1)*128, cb*64:(cb+1)*64]).max()
                alt = np.abs(got[rb*128:(rb+1)*128, cb*64:(cb+1)*64] - ref[rb*128:(rb+1)*128, (1-cb)*64:(2-cb)*64]).max()
                print('  rows', rb, 'cols', cb, 'err', blk, 'err_vs_swapped_cols', alt)

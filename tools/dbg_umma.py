import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfactor_b200 import _lib
K = int(sys.argv[1]); swap = int(sys.argv[2])
ctx = _lib.default_context()
rng = np.random.default_rng(0)
a = rng.integers(-4, 5, (128, K)).astype(np.float32)
b = rng.integers(-4, 5, (128, K)).astype(np.float32)
ref = a @ b.T
got = _lib.selftest_umma(ctx, torch.tensor(a).cuda(), torch.tensor(b).cuda(), bool(swap))
torch.cuda.synchronize()
got = got.cpu().numpy()
print('K', K, 'swap', swap, 'maxerr', float(np.abs(got - ref).max()), 'nnz', int((got != 0).sum()))
np.save(os.path.join(ROOT, 'gpurun_out', 'umma_K%d_s%d_got.npy' % (K, swap)), got)
np.save(os.path.join(ROOT, 'gpurun_out', 'umma_K%d_ref.npy' % K), ref)
